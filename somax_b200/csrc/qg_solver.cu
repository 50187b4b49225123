// PV inversion for the QG models: psi = ring0(Cm2l . Helm^-1_lambda . Cl2m . q).
//
// Replaces BaroclinicQG._invert_pv / finitevolx.pv_inversion(bc="dst") (reference
// qg/baroclinic.py:135-159; recipe SURVEY.md App. B.4).  The reference transforms the WHOLE
// (Ny, Nx) array it is given - ghost ring included - with a DST-I in both directions and divides
// by the 5-point eigenvalues (pinned by the reference's printed tutorial outputs, see
// oracle/elliptic.py), i.e. it solves the discrete system (delta_xx + delta_yy - lambda) psi = q
// on Ny x Nx unknowns with psi = 0 one cell outside the array EXACTLY.  Any exact direct solver
// of that system is therefore equivalent to rounding.  DST-I of length Nx = nx + 2 needs an FFT
// of length 2(nx+3), which for nx = 2^p is never a power of two (8195 = 5*11*149), so this
// solver is B200-first instead:
//
//   FFT path (nx = 2^p): array columns 1..nx-1 are transformed with a DST-I of size nx-1 (an
//     in-shared-memory complex FFT of length nx); columns 0, nx and nx+1 are border unknowns
//     g0, g1, g2 handled by a Schur complement (tools/proto_bordered3.py): solve 1 gives v, the
//     border right-hand sides f_0 - b v(1), f_nx - b v(nx-1), f_nx+1 go through a dense DST in y
//     and a host-inverted 3 x 3 matrix per y-mode, solve 2 adds the harmonic correction
//     b sin(pi k / nx) (g0 +- g1) (sign by the parity of the x-wavenumber k).
//   dense path (any nx <= 2046): x transform as a dense DST-I(nx+2) matrix product.
//   Both: the y direction is solved per x-wavenumber by the Thomas algorithm, marching in y
//     with one thread per wavenumber (coalesced), carried in fp64 even for fp32 data (the
//     second-difference recurrences lose ~ (ny/pi)^2 eps otherwise), coefficients from a
//     compact host-built fp64 table (the Toeplitz recurrence converges to a fixed point).
#include "qg_solver.cuh"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "fft.cuh"

namespace sb {

constexpr int QG_MAX_NL = 4;
constexpr int TH_COLS = 64;   // columns (threads) per Thomas CTA

struct Mix { double c[QG_MAX_NL][QG_MAX_NL]; };

struct QgSolver {
  // ny = SOLVER rows (every row of the array is an unknown: L.Ny for a whole grid, the owned rows
  // for the row stage of a slab); solver row j is field row jo + j; ylo / yhi: row 0 / ny-1 is a
  // physical ring row.  nx = interior columns (array columns = nx + 2).
  int dtype, batch, nl, ny, nx, kind;
  int jo = 0, ylo = 1, yhi = 1;
  double dx, dy;
  Layout L;
  int np, ncols, planes;
  void* S = nullptr; void* W = nullptr;
  double* ctab = nullptr; long long* coff = nullptr; int* krow = nullptr; double* cinf = nullptr;
  int kbad[QG_MAX_NL] = {0, 0, 0, 0}; int KB = 0, KBs = 2; double* dbad = nullptr; double* dbad1 = nullptr; double* meet1 = nullptr; void* part = nullptr;
  double* bsig = nullptr; void* sig2n = nullptr; double* minv = nullptr; double* sintab = nullptr;
  double* zpart = nullptr; unsigned* zcount = nullptr;   // border_dst: partial sums of term-split CTAs
  void* bext = nullptr;                             // [plane][3][ny]: border columns 0, nx+1 and (raw, forward only) nx
  // slab-distributed model: the row stage stores its spectral rows straight into the column arrays
  // of the ranks that own the strips (scatter), the last sweep stores its tiles straight into the
  // row arrays of the ranks that own the rows (push) - peer memory over NVLink, no transpose kernels
  void* sc_peer[16] = {nullptr}; int sc_lgspr = -1, sc_row0 = 0, sc_ny = 0;
  void* ps_peer[16] = {nullptr}; int ps_row0[17] = {0}; int ps_n = 0;
  // slab-distributed model: the border partial sums of a rank's strips are pre-summed on the rank
  // (pvec[rank][plane][2][ny], fp64) and only those vectors are exchanged
  double* pvec = nullptr; int pv_n = 0, pv_me = 0, pv_s0 = 0, pv_s1 = 0;
  double* rvec = nullptr; double* ghat = nullptr; double* gvec = nullptr; float* gvecf = nullptr; double* meet = nullptr; double* meetc = nullptr;
  void* tw = nullptr; void* twc = nullptr; void* twb = nullptr; void* dstmat = nullptr;
  FftPlan plan;
  Mix l2m, m2l;
  int nheavy = 0;
  int nseg = 1, seg_len = 0;                        // segmented sweeps (ThomasTab)
  bool seg_plain = true;                            // false: only the LOWK class is segmented
  double* segbuf = nullptr; double* segprod = nullptr;
  float* ckpt = nullptr; int nck = 0;               // E checkpoints of the PLAIN fp32 strips (ThomasTab)
  int* vwarm = nullptr; int vwarm_max = 0;
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  size_t bytes = 0;
};

// Spectral arrays (S, W) are stored BLOCKED in 64-column strips so that the y-sweeps can move a
// (rows x 64 columns) tile with a single bulk copy: within a plane, element (row j, column k)
// lives at ((k / 64) * ny + j) * 64 + (k % 64); the plane stride is ny * np, np = roundup(nx, 64).
constexpr int GS_SMALL_NY = 1024;   // up to here: one thread per output (border_gsolve_small)

// ------------------------------------------------------------------------------------------
// row kernels, FFT path, compile-time sized (the ones actually launched for nx = 2^LGN)
//   FWD: in = q (padded field), mix = Cl2m, out = S (spectral rows; slot n-1 = raw border column)
//   INV: in = S, mix = Cm2l (layer mixing commutes with the transform), out = psi (padded field)
// One CTA row-group per (member, y-row); the nl mixed rows are transformed one after another in
// the same shared-memory line.  Loads are 128-bit and batched per source layer.
// ------------------------------------------------------------------------------------------
template <int LGN> struct RowCfg {
  static constexpr int n = 1 << LGN;
  static constexpr int EPT = LGN >= 12 ? 16 : (LGN >= 8 ? 8 : (LGN >= 7 ? 4 : 2));
  static constexpr int G = n / EPT;                       // threads per row
  static constexpr int RPB = G >= 128 ? 1 : (G == 32 ? 8 : 128 / G);      // rows per block
  // G == 32: a row belongs to one warp, so the passes synchronise with __syncwarp only
  static constexpr bool WARP_ROWS = G == 32;
  static constexpr int threads = G * RPB;
  static constexpr int minblocks = threads >= 512 ? 2 : (threads >= 256 ? 3 : 4);
};

template <typename T>
struct RowArgsCT {
  Layout L;
  // ny = solver rows (the whole array: L.Ny; a slab: the rows it owns); solver row j is field row
  // jo + j.  ylo / yhi: solver row 0 / ny-1 is a physical ring row.  ringmode: FWD = the ring of the
  // input is to be read as zero (boundary condition applied on load); INV = the ring of psi is not
  // written (BaroclinicQG zeroes it, qg/baroclinic.py:157-158; it stays at its initial zero).
  // bext: [plane][3][ny] border columns 0 and nx+1 (raw mixed right-hand side in, solution out) and
  // the raw column nx (FWD).
  // Scatter (FWD, slab-distributed model, lgspr >= 0): spectral element (row j, slot p) goes to the
  // column array of the rank that owns strip p / 64, peer[(p >> 6) >> lgspr], at row prow0 + j of its
  // pny rows - a store into peer memory over NVLink - instead of this solver's own array.
  int ny, np, nl, nrows, jo, ylo, yhi, ringmode;
  T* bext;
  T* peer[16];
  int lgspr, prow0, pny;
  T mix[QG_MAX_NL][QG_MAX_NL];
  const C2<T>* tw;    // exp(-i pi t / n), t = 0..2n-1 (real-odd split)
  const C2<T>* twc;   // compact per-pass butterfly twiddles (fft.cuh: twc_offset)
  const C2<T>* twb;   // three-pass kernel: [LG1][G] pass-1 and [LG2][R3] pass-2 twiddle powers (or null)
  T scale;
};

template <typename T, int LGN, int G, int PASS>
__device__ __forceinline__ void fft_passes_ct(C2<T>* s, int lt, const C2<T>* __restrict__ tw, bool valid) {
  if constexpr (PASS < FftCT<LGN>::npass) {
    constexpr int lr = FftCT<LGN>::lgr(PASS);
    constexpr int LGLC = LGN - 3 * PASS;      // every earlier pass is radix 8
    if (valid) fft_dif_pass_ct<T, LGN, LGLC, (1 << lr), G>(s, lt, tw);
    if (RowCfg<LGN>::WARP_ROWS) __syncwarp(); else __syncthreads();
    fft_passes_ct<T, LGN, G, PASS + 1>(s, lt, tw, valid);
  }
}

// Destination of spectral element (row j, slot p) of plane `plane` in the forward transform.
template <typename T>
__device__ __forceinline__ T* fwd_dst(const RowArgsCT<T>& A, T* out, int plane, int j, int p) {
  if (A.lgspr >= 0)
    return A.peer[(p >> 6) >> A.lgspr] + (size_t)plane * A.pny * A.np + sp_off(A.pny, A.prow0 + j, p);
  return out + (size_t)plane * A.ny * A.np + sp_off(A.ny, j, p);
}

// Border columns 0 (which = 0) and nx+1 (which = 1) of one row, handled by two threads of the row:
// FWD mixes the layers' raw values into bext (zero when the input ring is read as zero); INV mixes
// the solved modal values of bext into psi unless the ring of psi is left alone.
template <typename T, bool INV>
__device__ __forceinline__ void row_border_cols(const RowArgsCT<T>& A, const T* __restrict__ in,
                                                T* __restrict__ out, int b, int a, int j, int fj,
                                                int which, int n) {
  if (A.ringmode) {
    if (!INV) A.bext[(((size_t)b * A.nl + a) * 3 + which) * A.ny + j] = T(0);
    return;
  }
  const int col = which ? n + 1 : 0;
  T acc = 0;
  for (int c = 0; c < A.nl; ++c) {
    const T v = INV ? A.bext[(((size_t)b * A.nl + c) * 3 + which) * A.ny + j]
                    : in[(((size_t)b * A.nl + c) * A.L.Ny + fj) * A.L.pitch + OFF + col];
    acc += A.mix[a][c] * v;
  }
  if (INV) out[(((size_t)b * A.nl + a) * A.L.Ny + fj) * A.L.pitch + OFF + col] = acc;
  else A.bext[(((size_t)b * A.nl + a) * 3 + which) * A.ny + j] = acc;
}

template <typename T, int LGN, bool INV>
__global__ void __launch_bounds__(RowCfg<LGN>::threads, RowCfg<LGN>::minblocks)
rowdst_fft_ct(RowArgsCT<T> A, const T* __restrict__ in, T* __restrict__ out) {
  using Cfg = RowCfg<LGN>;
  constexpr int n = Cfg::n, G = Cfg::G, EPT = Cfg::EPT;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lrow = threadIdx.x / G, lt = threadIdx.x % G;
  constexpr int plen = fft_padded_len(n);
  C2<T>* s = reinterpret_cast<C2<T>*>(smem_raw) + (size_t)lrow * plen;
  T* z = reinterpret_cast<T*>(s);
  const int row = blockIdx.x * Cfg::RPB + lrow;
  const bool valid = row < A.nrows;
  const int b = valid ? row / A.ny : 0, j = valid ? row - b * A.ny : 0;
  const int fj = A.jo + j;
  auto zi = [&](int t) { return 2 * fft_pad(t >> 1) + (t & 1); };
  // source row of layer/mode c and destination row of mode/layer a
  // element p of the source row of layer/mode c, and of the destination row of mode/layer a
  // (field rows are contiguous; spectral rows are blocked in 64-column strips)
  auto src = [&](int c, int p) -> const T* {
    return INV ? in + ((size_t)b * A.nl + c) * A.ny * A.np + sp_off(A.ny, j, p)
               : in + (((size_t)b * A.nl + c) * A.L.Ny + fj) * A.L.pitch + OFF + 1 + p;
  };
  auto dst = [&](int a, int p) -> T* {
    if (INV) return out + (((size_t)b * A.nl + a) * A.L.Ny + fj) * A.L.pitch + OFF + 1 + p;
    return fwd_dst<T>(A, out, b * A.nl + a, j, p);
  };
  auto raw_border = [&](int a) -> T* { return A.bext + (((size_t)b * A.nl + a) * 3 + 2) * A.ny + j; };
  const bool ring_row = valid && A.ringmode && ((j == 0 && A.ylo) || (j == A.ny - 1 && A.yhi));
  if (ring_row) {
    // FWD: a ring row of the boundary-conditioned input is zero, so is its transform;
    // INV: the ring of psi is left alone
    if (!INV)
      for (int a = 0; a < A.nl; ++a) {
        for (int p = lt; p < n; p += G) *dst(a, p) = T(0);
        if (lt < 3) A.bext[(((size_t)b * A.nl + a) * 3 + lt) * A.ny + j] = T(0);
      }
  }
  const bool work = valid && !ring_row;    // (block-wide barriers below are reached by everyone)

  for (int a = 0; a < A.nl; ++a) {
    if (work && lt < 2) row_border_cols<T, INV>(A, in, out, b, a, j, fj, lt, n);
    if (work) {
      if constexpr (EPT >= 4) {
        constexpr int NV = EPT / 4;
        Vec4<T> acc[NV];
#pragma unroll
        for (int e = 0; e < NV; ++e) acc[e] = Vec4<T>{0, 0, 0, 0};
        for (int c = 0; c < A.nl; ++c) {
          Vec4<T> v[NV];
#pragma unroll
          for (int e = 0; e < NV; ++e) v[e] = ld4(src(c, 4 * (lt + e * G)));
          const T mx = A.mix[a][c];
#pragma unroll
          for (int e = 0; e < NV; ++e) {
            acc[e].x += mx * v[e].x; acc[e].y += mx * v[e].y;
            acc[e].z += mx * v[e].z; acc[e].w += mx * v[e].w;
          }
        }
#pragma unroll
        for (int e = 0; e < NV; ++e) {
          const int t = 4 * (lt + e * G) + 1;     // x_t .. x_{t+3}
          const T vals[4] = {acc[e].x, acc[e].y, acc[e].z, acc[e].w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (t + u < n) { z[zi(t + u)] = vals[u]; z[zi(2 * n - t - u)] = -vals[u]; }
            else if (INV) *dst(a, n - 1) = vals[u];        // border column (x index n)
            else *raw_border(a) = vals[u];
          }
        }
      } else {
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          const int p = lt + e * G;
          T val = 0;
          for (int c = 0; c < A.nl; ++c) val += A.mix[a][c] * *src(c, p);
          const int t = p + 1;
          if (t < n) { z[zi(t)] = val; z[zi(2 * n - t)] = -val; }
          else if (INV) *dst(a, n - 1) = val;
          else *raw_border(a) = val;
        }
      }
      if (lt == 0) { z[zi(0)] = 0; z[zi(n)] = 0; }
    }
    if (Cfg::WARP_ROWS) __syncwarp(); else __syncthreads();
    fft_passes_ct<T, LGN, G, 0>(s, lt, A.twc, work);
    if (work) {
#pragma unroll
      for (int k0 = 1; k0 <= n / 2; k0 += G) {
        const int k = k0 + lt;
        if (k <= n / 2) {
          T Xk, Xnk;
          dst_split_ct<T, LGN>(s, k, A.tw, Xk, Xnk);
          *dst(a, k - 1) = INV ? A.scale * Xk : Xk;
          if (k != n - k) *dst(a, n - k - 1) = INV ? A.scale * Xnk : Xnk;
        }
      }
    }
    if (Cfg::WARP_ROWS) __syncwarp(); else __syncthreads();
  }
}

template <typename T, int LGN, bool INV>
static int launch_rowdst_ct(const RowArgsCT<T>& A, const T* in, T* out, cudaStream_t st) {
  using Cfg = RowCfg<LGN>;
  constexpr size_t smem = (size_t)fft_padded_len(Cfg::n) * sizeof(C2<T>) * Cfg::RPB;
  if (int rc = ensure_dyn_smem((const void*)rowdst_fft_ct<T, LGN, INV>, smem)) return rc;
  const int blocks = (A.nrows + Cfg::RPB - 1) / Cfg::RPB;
  prof_begin(INV ? "rowdst_inv_fft" : "rowdst_fwd_fft", st);
  rowdst_fft_ct<T, LGN, INV><<<blocks, Cfg::threads, smem, st>>>(A, in, out);
  SB_LAUNCH_CHECK();
  return 0;
}

// ------------------------------------------------------------------------------------------
// row kernels, FFT path, large rows (n = 4096 / 8192 / 16384, fp32): three radix-16/32 passes with
// the butterflies held in registers.  The shared-memory line is crossed 6 times per transform
// (write after pass 1, read + write in passes 2 and 3, read in the real-odd split) instead of 12
// for the radix-8 kernel above, which is bound by shared-memory wavefronts:
//   pass 1 (radix R1, stride G = n/R1): thread lt owns complex elements lt + G q, loaded straight
//     from global memory (layer mixing and the odd extension happen in registers: element k is
//     (x_2k, x_2k+1) for k < n/2 and -(x_2n-2k, x_2n-2k-1) above), result to s[q' G + lt];
//   pass 2 (radix R2 inside blocks of G): butterfly (b, pos) reads s[b G + pos + R3 q], writes the
//     transposed layout s[pos P2 + b R2 + q'], P2 = n/R3 + 1, so that
//   pass 3 (radix R3) reads s[q P2 + t3] with consecutive lanes on consecutive words, and writes
//     frequency k = b + R1 (q2' + R2 q3') in natural order at s[k + (k >> LG1)];
//   split: pairs (k, n-k) are read in natural order, conflict-free.
// Every layout change is a read-all / barrier / write-all on the same line.
// ------------------------------------------------------------------------------------------
template <int LGN> struct BigCfg {
  static constexpr int LG1 = LGN >= 13 ? 5 : 4;
  static constexpr int LG2 = LGN >= 14 ? 5 : 4;
  static constexpr int LG3 = LGN - LG1 - LG2;
  static_assert(LG3 == 4, "supported: n = 4096, 8192, 16384");
  static constexpr int n = 1 << LGN, R1 = 1 << LG1, R2 = 1 << LG2, R3 = 1 << LG3;
  static constexpr int G = n / R1;                 // threads per row = butterflies of pass 1
  static constexpr int B2 = R1 / R2, B3 = R1 / R3; // butterflies per thread in passes 2, 3
  static constexpr int P2 = n / R3 + 1;
  static constexpr int slen = (R3 * P2 > n + (n >> LG1) + 1) ? R3 * P2 : n + (n >> LG1) + 1;
  static constexpr int minblocks = LGN >= 14 ? 1 : 2;
  // PAIR (n = 8192, 16384): every thread runs a pass-3 butterfly AND its mirror (frequencies k and
  // n - k), so the real-odd split happens in registers and the results leave for global memory
  // straight from there: the line is crossed 6 times per transform instead of 8 (the kernel is
  // bound by shared-memory wavefronts).  Needs two butterflies per thread in
  // pass 3 and R1 = 32 = one warp of pass-1 digits.
  static constexpr bool PAIR = (B3 == 2) && (R1 == 32);
};

template <int LGN, bool INV>
__global__ void __launch_bounds__(BigCfg<LGN>::G, BigCfg<LGN>::minblocks)
rowdst_fft_big(RowArgsCT<float> A, const float* __restrict__ in, float* __restrict__ out) {
  using Cfg = BigCfg<LGN>;
  using C = C2<float>;
  constexpr int n = Cfg::n, G = Cfg::G, R1 = Cfg::R1, R2 = Cfg::R2, R3 = Cfg::R3;
  constexpr int LG1 = Cfg::LG1, LG2 = Cfg::LG2, LG3 = Cfg::LG3, P2 = Cfg::P2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* s = reinterpret_cast<C*>(smem_raw);
  const int lt = threadIdx.x, lane = lt & 31;
  const int row = blockIdx.x;
  const int b = row / A.ny, j = row - b * A.ny, fj = A.jo + j;
  // The butterfly twiddles of a thread are the same for every mode: they are copied to shared
  // memory once per CTA (behind the line) and re-read from there - the global loads in front of
  // passes 1, 2 and the split were 6 % of the kernel's stall samples (L1 / L2 latency no other warp
  // covers: all eight warps of the CTA are in the same phase).
  C* tw1 = s + ((Cfg::slen + 1) & ~1);             // [LG1][G]: exp(-2 pi i lt 2^jj / n); 16-byte aligned
  C* tw2 = tw1 + LG1 * G;                          // [LG2][R3]: exp(-2 pi i pos 2^jj / G)
  C* tw0 = tw2 + LG2 * R3;                         // [G]: split twiddle exp(-i pi lt / n)
  if (A.ringmode && ((j == 0 && A.ylo) || (j == A.ny - 1 && A.yhi))) {
    // FWD: a ring row of the boundary-conditioned input is zero, so is its transform;
    // INV: the ring of psi is left alone
    if (!INV)
      for (int a = 0; a < A.nl; ++a) {
        for (int p = lt; p < n; p += G) *fwd_dst<float>(A, out, b * A.nl + a, j, p) = 0.f;
        if (lt < 3) A.bext[(((size_t)b * A.nl + a) * 3 + lt) * A.ny + j] = 0.f;
      }
    return;
  }

  // The rows of the wave after this one (2 CTAs on each of 148 SMs, dispatched in row order) are
  // pulled into L2 now, so that their first mode starts from an L2 hit instead of a DRAM miss.
  {
    const int prow = row + 2 * 148;
    if (prow < A.nrows) {
      const int pb = prow / A.ny, pj = prow - pb * A.ny;
      if (!INV) {
        if (lt < A.nl) {
          const float* src = in + (((size_t)pb * A.nl + lt) * A.L.Ny + (A.jo + pj)) * A.L.pitch + OFF + 1;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src), "r"((unsigned)(n * sizeof(float))) : "memory");
        }
      } else {
        // a spectral row is n / 64 segments of 256 bytes per mode: two 128-byte lines each
        for (int e = lt; e < A.nl * (n / SP_W) * 2; e += G) {
          const int c = e / ((n / SP_W) * 2), r2 = e - c * ((n / SP_W) * 2);
          const float* src = in + ((size_t)pb * A.nl + c) * A.ny * A.np + sp_off(A.ny, pj, (r2 >> 1) * SP_W) + (r2 & 1) * 32;
          asm volatile("prefetch.global.L2 [%0];\n" ::"l"(src) : "memory");
        }
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < LG1; ++jj) tw1[jj * G + lt] = A.twb[jj * G + lt];
  if (lt < LG2 * R3) tw2[lt] = A.twb[LG1 * G + lt];
  tw0[lt] = A.tw[lt];
  // (visible after the barrier that follows the staging of the first mode)

  for (int a = 0; a < A.nl; ++a) {
    if (lt < 2) row_border_cols<float, INV>(A, in, out, b, a, j, fj, lt, n);
    // ---- stage the mixed, odd-extended row: z_t at word t of the line (128-bit, conflict-free).
    // A thread's vector holds x_{4i+1..4i+4}; the lower half needs x_{4i..4i+3}: x_4i comes from
    // the lane below by shuffle, and across a warp boundary lane 31 stores it for its neighbour
    // (lane 0 then writes only its three own words).  The mirrored upper half is the own vector
    // reversed and negated, which is aligned as it is.
    {
      float* z = reinterpret_cast<float*>(s);
      constexpr int NV = n / 4 / G;
      Vec4<float> acc[NV];
#pragma unroll
      for (int e = 0; e < NV; ++e) acc[e] = Vec4<float>{0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < A.nl; ++c) {
        const float mx = A.mix[a][c];
        // element p = 4 lt + 4 G e: field rows are contiguous, spectral rows advance 4G/64 strips
        const float* src = INV ? in + ((size_t)b * A.nl + c) * A.ny * A.np + sp_off(A.ny, j, 4 * lt)
                               : in + (((size_t)b * A.nl + c) * A.L.Ny + fj) * A.L.pitch + OFF + 1 + 4 * lt;
        const size_t estride = INV ? (size_t)(4 * G / SP_W) * A.ny * SP_W : (size_t)4 * G;
        Vec4<float> w[NV];
#pragma unroll
        for (int e = 0; e < NV; ++e) w[e] = ld4(src + e * estride);
#pragma unroll
        for (int e = 0; e < NV; ++e) {
          acc[e].x = fmaf(mx, w[e].x, acc[e].x); acc[e].y = fmaf(mx, w[e].y, acc[e].y);
          acc[e].z = fmaf(mx, w[e].z, acc[e].z); acc[e].w = fmaf(mx, w[e].w, acc[e].w);
        }
      }
#pragma unroll
      for (int e = 0; e < NV; ++e) {
        const int i = lt + e * G;
        const float lo = __shfl_up_sync(0xffffffffu, acc[e].w, 1);
        float hi = -acc[e].w;
        if (i == n / 4 - 1) {
          // x_n is the border column, not part of the transform: z_n = 0
          *(INV ? out + (((size_t)b * A.nl + a) * A.L.Ny + fj) * A.L.pitch + OFF + n
                : A.bext + (((size_t)b * A.nl + a) * 3 + 2) * A.ny + j) = acc[e].w;
          hi = 0.f;
        } else if (lane == 31) {
          z[4 * i + 4] = acc[e].w;           // x_{4(i+1)} for lane 0 of the next warp
        }
        if (lane == 0) {
          if (i == 0) z[0] = 0.f;            // z_0
          z[4 * i + 1] = acc[e].x; z[4 * i + 2] = acc[e].y; z[4 * i + 3] = acc[e].z;
        } else {
          st4(z + 4 * i, Vec4<float>{lo, acc[e].x, acc[e].y, acc[e].z});
        }
        st4(z + 2 * n - 4 * i - 4, Vec4<float>{hi, -acc[e].z, -acc[e].y, -acc[e].x});
      }
    }
    __syncthreads();
    // ---- pass 1
    C v[R1];
#pragma unroll
    for (int q = 0; q < R1; ++q) v[q] = s[lt + G * q];
    __syncthreads();
    fft_reg<float, R1>(v);
    {
      C wp[LG1];
#pragma unroll
      for (int jj = 0; jj < LG1; ++jj) wp[jj] = tw1[jj * G + lt];
      fft_reg_twiddle<float, R1>(v, wp);
    }
#pragma unroll
    for (int q = 0; q < R1; ++q) s[q * G + lt] = v[fft_reg_pos<R1>(q)];
    __syncthreads();

    // ---- pass 2
    {
      C u[Cfg::B2][R2];
#pragma unroll
      for (int i = 0; i < Cfg::B2; ++i) {
        const int t2 = lt + G * i, b2 = t2 >> LG3, pos = t2 & (R3 - 1);
#pragma unroll
        for (int q = 0; q < R2; ++q) u[i][q] = s[b2 * G + pos + R3 * q];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < Cfg::B2; ++i) {
        const int t2 = lt + G * i, b2 = t2 >> LG3, pos = t2 & (R3 - 1);
        fft_reg<float, R2>(u[i]);
        C wp[LG2];
#pragma unroll
        for (int jj = 0; jj < LG2; ++jj) wp[jj] = tw2[jj * R3 + pos];
        fft_reg_twiddle<float, R2>(u[i], wp);
#pragma unroll
        for (int q = 0; q < R2; ++q)
          s[pos * P2 + (Cfg::PAIR ? q * R1 + b2 : b2 * R2 + q)] = u[i][fft_reg_pos<R2>(q)];
      }
    }
    __syncthreads();

    // destination of X_k (element p = k - 1 of the spectral row / field column k of the psi row)
    auto store_out = [&](int k, float val) {
      if (INV) out[(((size_t)b * A.nl + a) * A.L.Ny + fj) * A.L.pitch + OFF + k] = A.scale * val;
      else *fwd_dst<float>(A, out, b * A.nl + a, j, k - 1) = val;
    };
    auto split_pair = [&](int k, const C& Ak, const C& Bk, const C& wk) {
      const C E = {0.5f * (Ak.x + Bk.x), 0.5f * (Ak.y - Bk.y)};
      const C O = {0.5f * (Ak.y + Bk.y), -0.5f * (Ak.x - Bk.x)};
      const C wO = cmul(wk, O);
      store_out(k, -0.5f * (E.y + wO.y));
      store_out(n - k, 0.5f * (E.y - wO.y));
    };
    if constexpr (Cfg::PAIR) {
      // ---- pass 3 on a butterfly (b3, q2) = (lane, warp) and its mirror (32 - lane, R2 - 1 - warp):
      // outputs k = b3 + R1 (q2 + R2 q) and n - k = b3' + R1 (q2' + R2 (R3 - 1 - q)).  Pass 2 stored
      // its output with b fastest, so both reads are conflict-free (consecutive / reversed words).
      const int w = lt >> 5, l = lt & 31;
      const int cola = w * R1 + l, colb = (R2 - 1 - w) * R1 + ((R1 - l) & (R1 - 1));
      C ua[R3], ub[R3];
#pragma unroll
      for (int q = 0; q < R3; ++q) { ua[q] = s[q * P2 + cola]; ub[q] = s[q * P2 + colb]; }
      __syncthreads();
      fft_reg<float, R3>(ua);
      fft_reg<float, R3>(ub);
      // lane 0 holds the frequencies k = R1 m, m = q2 + R2 q, whose mirrors R1 (R2 R3 - m) sit in
      // lane 0 of OTHER warps: those M = R2 R3 values meet in shared memory
      constexpr int M = R2 * R3;
      if (l == 0) {
#pragma unroll
        for (int q = 0; q < R3; ++q) {
          s[w + R2 * q] = ua[fft_reg_pos<R3>(q)];
          s[(R2 - 1 - w) + R2 * q] = ub[fft_reg_pos<R3>(q)];
        }
      } else {
        constexpr int KB = 8;
        // X_k of consecutive q sit a constant stride apart (R1 R2 columns = R1 R2 / 64 strips).  When
        // the row is scattered over the strip owners of the slab model the element keeps its offset
        // inside the owner's (full-size) array and only the base pointer changes with the strip.
        const bool scatter = !INV && A.lgspr >= 0;
        const int kk0 = l + R1 * w;
        constexpr int QSTRIPS = R1 * R2 / SP_W;
        const int sk = (kk0 - 1) >> 6, snk = (n - kk0 - 1) >> 6;
        const size_t offk = scatter ? ((size_t)b * A.nl + a) * A.pny * A.np + sp_off(A.pny, A.prow0 + j, kk0 - 1) : 0;
        const size_t offnk = scatter ? ((size_t)b * A.nl + a) * A.pny * A.np + sp_off(A.pny, A.prow0 + j, n - kk0 - 1) : 0;
        const ptrdiff_t sstride = (ptrdiff_t)QSTRIPS * A.pny * SP_W;
        float* pk = INV ? out + (((size_t)b * A.nl + a) * A.L.Ny + fj) * A.L.pitch + OFF + kk0
                        : out + ((size_t)b * A.nl + a) * A.ny * A.np + sp_off(A.ny, j, kk0 - 1);
        float* pnk = INV ? out + (((size_t)b * A.nl + a) * A.L.Ny + fj) * A.L.pitch + OFF + (n - kk0)
                         : out + ((size_t)b * A.nl + a) * A.ny * A.np + sp_off(A.ny, j, n - kk0 - 1);
        const ptrdiff_t dstride = INV ? (ptrdiff_t)(R1 * R2) : (ptrdiff_t)(R1 * R2 / SP_W) * A.ny * SP_W;
        // split twiddles w_k = exp(-i pi k / n), k = kk0 + R1 R2 q: one load, then the 16 rotations
        // exp(-i pi q / R3) are compile-time constants (the table load per k was 10 % of the
        // kernel's LSU wavefronts, which bound it)
        static_assert(R3 == 16, "rotation table");
        constexpr float RC[16] = {1.f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
                                  0.70710678118654752f, 0.55557023301960218f, 0.38268343236508978f,
                                  0.19509032201612825f, 0.f, -0.19509032201612825f, -0.38268343236508978f,
                                  -0.55557023301960218f, -0.70710678118654752f, -0.83146961230254524f,
                                  -0.92387953251128674f, -0.98078528040323043f};
        constexpr float RS[16] = {0.f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
                                  0.70710678118654752f, 0.83146961230254524f, 0.92387953251128674f,
                                  0.98078528040323043f, 1.f, 0.98078528040323043f, 0.92387953251128674f,
                                  0.83146961230254524f, 0.70710678118654752f, 0.55557023301960218f,
                                  0.38268343236508978f, 0.19509032201612825f};
        const C w0 = tw0[kk0];
        const float qs = INV ? 0.25f * A.scale : 0.25f;
#pragma unroll
        for (int q0 = 0; q0 < R3; q0 += KB) {
          C wk[KB];
#pragma unroll
          for (int u = 0; u < KB; ++u) {     // w0 * (RC - i RS)
            const int q = q0 + u;
            wk[u] = cmul_cs(w0, RC[q], RS[q]);
          }
#pragma unroll
          for (int u = 0; u < KB; ++u) {
            const int q = q0 + u;
            // (X_k, X_{n-k}) = (-(e + o) / 4, (e - o) / 4), e = A.y - B.y, o = w.y (A.y + B.y) - w.x (A.x - B.x)
            const C Ak = ua[fft_reg_pos<R3>(q)], Bk = ub[fft_reg_pos<R3>(R3 - 1 - q)];
            const float2 p = __fadd2_rn(make_float2(Ak.y, Ak.y), make_float2(-Bk.y, Bk.y));
            const float o = fmaf(wk[u].y, p.y, -(wk[u].x * (Ak.x - Bk.x)));
            const float2 X = __fmul2_rn(__fadd2_rn(make_float2(p.x, p.x), make_float2(o, -o)), make_float2(-qs, qs));
            if (!scatter) {
              pk[q * dstride] = X.x;
              pnk[-q * dstride] = X.y;
            } else {
              A.peer[(sk + QSTRIPS * q) >> A.lgspr][offk + q * sstride] = X.x;
              A.peer[(snk - QSTRIPS * q) >> A.lgspr][offnk - q * sstride] = X.y;
            }
          }
        }
      }
      __syncthreads();
      if (lt < M / 2) {      // pairs (R1 m, R1 (M - m)), m = 1 .. M/2 (m = M/2 pairs with itself: both stores agree)
        const int m = lt + 1, k = R1 * m;
        split_pair(k, s[m], s[M - m], A.tw[k]);
      }
    } else {
    // ---- pass 3
    {
      C u[Cfg::B3][R3];
#pragma unroll
      for (int i = 0; i < Cfg::B3; ++i) {
        const int t3 = lt + G * i;
#pragma unroll
        for (int q = 0; q < R3; ++q) u[i][q] = s[q * P2 + t3];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < Cfg::B3; ++i) {
        const int t3 = lt + G * i, b3 = t3 >> LG2, q2 = t3 & (R2 - 1);
        fft_reg<float, R3>(u[i]);
#pragma unroll
        for (int q = 0; q < R3; ++q) {
          const int k = b3 + R1 * (q2 + R2 * q);
          s[k + (k >> LG1)] = u[i][fft_reg_pos<R3>(q)];
        }
      }
    }
    __syncthreads();

    // ---- real-odd split, pairs (k, n-k), k = 1 + lt + G m, read from the line in natural order
    {
      const int k1 = 1 + lt;
      constexpr int NK = (n / 2) / G;      // k = n/2 (thread G-1, last m) pairs with itself: both stores agree
      constexpr int KB = 8;
#pragma unroll 1
      for (int m0 = 0; m0 < NK; m0 += KB) {
        C w[KB];
#pragma unroll
        for (int u = 0; u < KB; ++u) w[u] = A.tw[k1 + (m0 + u) * G];
#pragma unroll
        for (int u = 0; u < KB; ++u) {
          const int k = k1 + (m0 + u) * G;
          split_pair(k, s[k + (k >> LG1)], s[(n - k) + ((n - k) >> LG1)], w[u]);
        }
      }
    }
    }
    __syncthreads();
  }
}

template <int LGN, bool INV>
static int launch_rowdst_big(const RowArgsCT<float>& A, const float* in, float* out, cudaStream_t st) {
  using Cfg = BigCfg<LGN>;
  constexpr size_t smem = (size_t)(((Cfg::slen + 1) & ~1) + Cfg::LG1 * Cfg::G + Cfg::LG2 * Cfg::R3 + Cfg::G) * sizeof(C2<float>);
  if (int rc = ensure_dyn_smem((const void*)rowdst_fft_big<LGN, INV>, smem)) return rc;
  prof_begin(INV ? "rowdst_inv_fft" : "rowdst_fwd_fft", st);
  rowdst_fft_big<LGN, INV><<<A.nrows, Cfg::G, smem, st>>>(A, in, out);
  SB_LAUNCH_CHECK();
  return 0;
}

template <typename T, bool INV>
static int launch_rowdst(int lgn, const RowArgsCT<T>& A, const T* in, T* out, cudaStream_t st) {
  if constexpr (sizeof(T) == 4) {
    if (A.twb) {
      if (lgn == 12) return launch_rowdst_big<12, INV>(A, in, out, st);
      if (lgn == 13) return launch_rowdst_big<13, INV>(A, in, out, st);
      if (lgn == 14) return launch_rowdst_big<14, INV>(A, in, out, st);
    }
  }
  switch (lgn) {
#define SB_ROW_CASE(L) case L: return launch_rowdst_ct<T, L, INV>(A, in, out, st);
    SB_ROW_CASE(3) SB_ROW_CASE(4) SB_ROW_CASE(5) SB_ROW_CASE(6) SB_ROW_CASE(7) SB_ROW_CASE(8)
    SB_ROW_CASE(9) SB_ROW_CASE(10) SB_ROW_CASE(11) SB_ROW_CASE(12) SB_ROW_CASE(13)
#undef SB_ROW_CASE
    case 14:
      if constexpr (sizeof(T) == 4) return launch_rowdst_ct<T, 14, INV>(A, in, out, st);
    default:
      return fail(SOMAX_B200_ERR_UNSUPPORTED, "FFT solver: unsupported nx");
  }
}

// ------------------------------------------------------------------------------------------
// row kernels, dense path: out[k] = scale * sum_i Smat[i*n + k] * (mixed row)[i] over ALL n = Nx
// columns of the array (no border columns); ringmode as in RowArgsCT.
// ------------------------------------------------------------------------------------------
template <typename T, bool INV>
__global__ void rowdst_dense(Layout L, int ny, int n, int np, int nl, int ringmode, Mix mix,
                             const T* __restrict__ Smat, const T* __restrict__ in,
                             T* __restrict__ out, double scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw);   // nl * n
  const int row = blockIdx.x;
  const int b = row / ny, j = row - b * ny;
  const bool ring_row = ringmode && (j == 0 || j == ny - 1);
  if (ring_row) {
    if (!INV)
      for (int e = threadIdx.x; e < nl * n; e += blockDim.x) {
        const int a = e / n, k = e - a * n;
        out[((size_t)b * nl + a) * ny * np + sp_off(ny, j, k)] = T(0);
      }
    return;
  }
  for (int e = threadIdx.x; e < nl * n; e += blockDim.x) {
    int a = e / n, i = e - a * n;
    double val = 0;
    if (INV || !ringmode || (i > 0 && i < n - 1))
      for (int c = 0; c < nl; ++c) {
        T src = INV ? in[((size_t)b * nl + c) * ny * np + sp_off(ny, j, i)]
                    : in[(((size_t)b * nl + c) * L.Ny + j) * L.pitch + OFF + i];
        val += mix.c[a][c] * (double)src;
      }
    xs[e] = (T)val;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double acc[QG_MAX_NL] = {0, 0, 0, 0};
    for (int i = 0; i < n; ++i) {
      double sv = (double)Smat[(size_t)i * n + k];
#pragma unroll
      for (int a = 0; a < QG_MAX_NL; ++a)
        if (a < nl) acc[a] += sv * (double)xs[a * n + i];
    }
    for (int a = 0; a < nl; ++a) {
      if (INV) {
        if (!ringmode || (k > 0 && k < n - 1))
          out[(((size_t)b * nl + a) * L.Ny + j) * L.pitch + OFF + k] = (T)(scale * acc[a]);
      } else {
        out[((size_t)b * nl + a) * ny * np + sp_off(ny, j, k)] = (T)(scale * acc[a]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Thomas sweeps along y, one thread per x-wavenumber.
//   normalised system per column: x_{j-1} + delta x_j + x_{j+1} = dy^2 f_j
//   c_j = 1/(delta - c_{j-1}),  d_j = (dy^2 f_j - d_{j-1}) c_j,  x_j = d_j - c_j x_{j+1}
// c_j is a property of the grid only: it is tabulated on the host in fp64 and converges to a
// fixed point after J rows (J << ny for almost every wavenumber), so rows below J read a
// per-strip table and the rest use the constant.  With exact coefficients the recurrence error
// of a column grows like (n / k pi) eps: the few low-k strips carry the recurrence in fp64 in
// both pipelines, everything else runs in the pipeline's own precision.
// ------------------------------------------------------------------------------------------
struct ThomasTab {
  // Coefficient table blocked like the data: for (mode m, strip) the rows i = 0..Js-1 of the
  // recurrence c_i for the strip's 64 columns, contiguous ([i][64] doubles).  Rows >= Js use cinf.
  const double* ctabB; const long long* tabOff; const int* Jstrip;
  const double* cinf;        // [m][np] fixed point per column
  const double* meetc;       // [m][2][np]: c[m1-1] and c[ny-1-m1] per column (meeting-point solve)
  int kbad[QG_MAX_NL]; int KB;
  double* dbad; double* dbad1;   // fp64 side buffers [plane][j][KB] of the indefinite columns (solve 1 / 2)
  double* meet; double* meet1;   // [plane][2][np]: last eliminated value of each half (solve 1 / 2)
  void* part;                    // [plane][2: even / odd column][2 nstrip][ny]: per-warp partial border sums (KIND 1)
  const void* sig2n;             // [np] border weights (2/n) sin(pi k / n) in working precision, zero-padded
  int ny, np, ncols, nl, nstrip;
  double dy2;
  // Segmented sweeps (slab-distributed model, where a rank owns few strips and a sweep CTA is a
  // bare serial chain): each half-column is cut into nseg segments of seg_len rows handled by
  // different CTAs.  The recurrence y_s = g_s - c_s y_{s-1} is linear, so pass 1 (probe) runs every
  // segment from a zero carry without storing and records its end value e; pass 2 (apply) starts
  // segment k from cin_k = e_{k-1} + p_{k-1} cin_{k-1} (p = product of -c_s over a segment,
  // tabulated on the host) and does the real work.  pass 0 = unsegmented (nseg = 1).
  int nseg, seg_len, pass;
  bool seg_plain;
  double* segbuf;            // [plane][half][seg][np] end values of pass 1
  const double* segprod;     // [kind: 0 elimination, 1 substitution][half][mode][seg][np]
  // PLAIN fp32 strips never store the eliminated border right-hand side E of the second solve:
  // thomas_vec_ckpt keeps one value of the recurrence per column and 16-row block and the KIND 2
  // substitution recomputes E block by block (the recurrence needs only the shared vector g and the
  // tabulated coefficients).  ckpt: [plane][strip][half][nck][64] floats, block m covers the
  // elimination rows [cnt - (m+1) 16, cnt - m 16) of the half and holds the value before its first.
  float* ckpt; int nck;
  const int* vwarm;          // [mode][strip] warm-up rows of a thomas_vec_ckpt segment (multiple of TH_RT)
  // Push (slab-distributed model, KIND 2 only): the finished tiles go straight into the row arrays
  // of the ranks that own the rows - rank r holds rows [prow[r], prow[r+1]) as
  // [plane][strip][prow[r+1] - prow[r]][64] at ppeer[r] (peer memory over NVLink) - instead of `out`.
  void* ppeer[16]; int prow[17]; int pn;
};


// Two-way ("burn at both ends") Thomas: every column is handled by TWO threads (blockIdx.z):
// half 0 eliminates rows 0..m1-1 downwards, half 1 eliminates rows ny-1..m1 upwards with the
// mirrored recurrence (same coefficient table by symmetry of the Toeplitz system).  The halves
// meet between rows m1-1 and m1 with a 2x2 solve, then substitute outwards.  Twice the
// parallelism of the classic sweep for the same memory passes.
//   SUBST = false: elimination  d_s = (dy^2 f_s - d_{s-1}) c_s           (s counts from the end)
//   SUBST = true : substitution x_s = d_s - c_s x_{s'}  outwards from the meeting point
// FROM_VEC (elimination only): right-hand side is gvec[plane][c & 1][j] for column c (border
// solve: (g0 + g1)(y) for the odd x-wavenumbers k = c + 1, (g0 - g1)(y) for the even ones).
// KIND (substitution only): 0 = plain solve, x is stored.  The bordered solver never stores the
// first solve's x: KIND 1 substitutes the eliminated first right-hand side D only to emit the
// weighted row sums sum_c sig2n[c] x[j][c] the border system needs, separately over the even
// and the odd columns (per-warp partials, reduced by border_reduce: their sum is v at block
// column 1, their difference v at block column n-1); KIND 2 substitutes D - bsig[c] * E (E =
// eliminated border right-hand side of the second solve, linearity of the elimination) and
// stores the final x.  7 instead of 8 array transfers per inversion and no separate dot pass.
// ---- bulk-async (TMA engine) helpers: 1-D cp.async.bulk + mbarrier, no tensor map needed ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem, const void* gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem, const void* smem, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n"
               ::"l"(gmem), "r"(smem_u32(smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

#if defined(SB_TH_DEBUG) || defined(SB_TH_PHASES)
__device__ unsigned long long g_dbg[4096 * 4];
__device__ unsigned g_dbg_n;
#endif
constexpr int TH_RT = 16;   // rows per register block of the recurrence

// Launch classes of a sweep.  LOWK (TAB = true): the few strips of near-singular low x-wavenumbers
// (and late-converging coefficient recurrences): fp64 carried recurrence for both pipelines,
// coefficient tiles staged through shared memory, big staged tiles because only a handful of
// CTAs exist and each is a serial chain over ny/2 rows.  PLAIN (TAB = false): everything else, in
// the pipeline's own precision, coefficients (needed for the first rows only) read from L2.
template <typename T, int KIND, bool TAB> struct ThRows {      // rows per staged tile
  // (LOWK fp32 used 64-row tiles when a sweep was a handful of unsegmented CTAs; with 8 segments there
  // are 240 of them and a 150 KB ring left room for only three PLAIN CTAs beside it on an SM, so the
  // PLAIN launch spilled into a second wave)
  static constexpr int v = TAB ? 32 : TH_RT;
};
template <typename T, int KIND, bool TAB> struct ThStages {   // ring depth
  // PLAIN rings are kept small enough that a whole sweep (123 strips x 3 planes x 2 halves at
  // 8192^2) is resident in ONE wave next to the LOWK CTAs: every CTA runs for the whole sweep, so
  // a second wave of a few CTAs would double the kernel time
  static constexpr int v = TAB ? 3 : (sizeof(T) == 4 ? 6 : (KIND == 2 ? 2 : 3));
};

// Plain fp32 recurrence over one register block, for the well-conditioned strips of the fp32
// pipeline: for x-wavenumbers k >~ n/32 the fp32 recurrence error stays at the rounding level
// (~1e-7 per column; the error of a column grows like (n / k pi) eps with exact coefficients),
// only the low-k strips need the fp64 carry.  Coefficient rows are plain floats.
template <bool SUBST, bool FROM_VEC, int KIND, bool UP, int MODE>
__device__ __forceinline__ void thomas_tile_f32(float (*A)[TH_COLS], const float (*Vt)[TH_COLS],
                                                const float (*Ct)[TH_COLS], const float* __restrict__ gv,
                                                int tid, int nr, int s0, int ilo, int Js, int cnt, int jb,
                                                bool act, float cfix, float kfix, float dy2, float bs,
                                                float ck, float& carry) {
  constexpr int dj = UP ? 1 : -1;
  float f[TH_RT], cj[TH_RT];
#pragma unroll
  for (int r = 0; r < TH_RT; ++r) {
    const bool ok = MODE != 2 || r < nr;
    const int rm = UP ? r : (MODE != 2 ? TH_RT - 1 - r : nr - 1 - r);
    const int i = SUBST ? cnt - 1 - (s0 + r) : s0 + r;
    if (MODE == 0) cj[r] = cfix;
    else if (MODE == 1) cj[r] = Ct[i - ilo][tid];
    else cj[r] = (ok && i < Js) ? Ct[i - ilo][tid] : cfix;
    if (FROM_VEC) f[r] = ok ? gv[jb + dj * r] : 0.f;
    else f[r] = ok ? A[ok ? rm : 0][tid] : 0.f;
  }
  if (KIND == 2) {
    // E of this block, recomputed: the elimination ran over the rows in the opposite order
    // (r = nr-1 .. 0), from the checkpointed value before the block's first row; same arithmetic
    // as thomas_vec_ckpt, so the values are the ones that kernel went through
    (void)Vt;
    float e = ck;
#pragma unroll
    for (int r = TH_RT - 1; r >= 0; --r) {
      if (MODE != 2 || r < nr) {
        const int i = cnt - 1 - (s0 + r);
        const float kc = (MODE == 0 || (MODE == 2 && i >= Js)) ? kfix : cj[r] * dy2;
        e = fmaf(-cj[r], e, kc * gv[jb + dj * r]);
        f[r] = fmaf(-bs, e, f[r]);
      }
    }
  }
  if (!SUBST) {
#pragma unroll
    for (int r = 0; r < TH_RT; ++r) f[r] = (MODE == 0 ? kfix : cj[r] * dy2) * f[r];
  }
#pragma unroll
  for (int r = 0; r < TH_RT; ++r)
    if (MODE != 2 || r < nr) { carry = fmaf(-cj[r], carry, f[r]); f[r] = carry; }
#pragma unroll
  for (int r = 0; r < TH_RT; ++r) {
    if (MODE != 2 || r < nr) {
      const int rm = UP ? r : (MODE != 2 ? TH_RT - 1 - r : nr - 1 - r);
      if (act) A[rm][tid] = f[r];
      else if (FROM_VEC) A[rm][tid] = 0.f;
    }
  }
}

// fp64 recurrence over one register block for one column (thread).  UP: memory rows ascend with
// the sequence (dj > 0).  MODE 0: full block, constant coefficient; MODE 1: full block, every row
// tabulated (coefficients in Ct); MODE 2: generic (ragged end, block straddling the convergence
// row, indefinite columns with their fp64 side buffer).  Phases are explicit (all loads, the
// carried chain, all stores) so the shared-memory accesses pipeline.
template <typename T, bool SUBST, bool FROM_VEC, int KIND, bool UP, int MODE>
__device__ __forceinline__ void thomas_tile(T (*A)[TH_COLS], const T (*Vt)[TH_COLS],
                                            const double (*Ct)[TH_COLS], const double* __restrict__ gv,
                                            double* __restrict__ Dt, const double* __restrict__ Dt1,
                                            int KB, int tid, int nr, int s0,
                                            int ilo, int Js, int cnt, int jb, bool act, bool bad,
                                            double cfix, double dy2, double bs, double& carry) {
  constexpr int dj = UP ? 1 : -1;
  double f[TH_RT], cj[TH_RT];
#pragma unroll
  for (int r = 0; r < TH_RT; ++r) {
    const bool ok = MODE != 2 || r < nr;
    const int rm = UP ? r : (MODE != 2 ? TH_RT - 1 - r : nr - 1 - r);   // memory row in the block
    const int i = SUBST ? cnt - 1 - (s0 + r) : s0 + r;                  // table index of the row
    if (MODE == 0) cj[r] = cfix;
    else if (MODE == 1) cj[r] = Ct[i - ilo][tid];
    else cj[r] = (ok && i < Js) ? Ct[i - ilo][tid] : cfix;
    if (FROM_VEC) f[r] = ok ? gv[jb + dj * r] : 0.0;
    else f[r] = ok ? (double)A[ok ? rm : 0][tid] : 0.0;
    if (KIND == 2) f[r] = fma(-bs, ok ? (double)Vt[ok ? rm : 0][tid] : 0.0, f[r]);
  }
  if (MODE != 0 && SUBST && bad) {
    // indefinite column: the eliminated right-hand side comes from the staged fp64 side tile
#pragma unroll
    for (int r = 0; r < TH_RT; ++r) {
      const int rm = UP ? r : (MODE != 2 ? TH_RT - 1 - r : nr - 1 - r);
      if (MODE != 2 || r < nr) f[r] = KIND == 2 ? fma(-bs, Dt1[rm * KB], Dt[rm * KB]) : Dt[rm * KB];
    }
  }
  if (!SUBST) {
    if (MODE == 0) {
      const double kf = cfix * dy2;
#pragma unroll
      for (int r = 0; r < TH_RT; ++r) f[r] = kf * f[r];
    } else {
#pragma unroll
      for (int r = 0; r < TH_RT; ++r) f[r] = (cj[r] * dy2) * f[r];
    }
  }
#pragma unroll
  for (int r = 0; r < TH_RT; ++r)
    if (MODE != 2 || r < nr) { carry = fma(-cj[r], carry, f[r]); f[r] = carry; }
#pragma unroll
  for (int r = 0; r < TH_RT; ++r) {
    if (MODE != 2 || r < nr) {
      const int rm = UP ? r : (MODE != 2 ? TH_RT - 1 - r : nr - 1 - r);
      if (act) A[rm][tid] = (T)f[r];
      else if (FROM_VEC) A[rm][tid] = T(0);
      if (MODE != 0 && !SUBST && bad) Dt[rm * KB] = f[r];
    }
  }
}

// y-sweep over staged tiles.  Each CTA owns one 64-column strip of one plane and one half
// (two-way elimination); it streams contiguous RT-row tiles of the strip through an NS-stage
// shared-memory ring: ONE cp.async.bulk per tile (TMA engine, completes on an mbarrier), the
// recurrence runs on shared memory only (register blocks of TH_RT rows), and the finished tile
// leaves with one bulk store.
// Threads 0..63 own one column each (two warps); a third warp's lane 0 is the PRODUCER: it
// issues every bulk load / store and waits for the stores' shared-memory reads, so those waits are
// off the recurrence warps' critical path (a sweep CTA is a serial chain over ny/2 rows: with the
// issue work in warp 0, a rank of the slab-distributed model that owns a few strips - one CTA per
// SM - spent most of each tile period in it).
#ifndef SB_TH_PRODUCER
#define SB_TH_PRODUCER 1
#endif
constexpr int TH_THREADS = TH_COLS + (SB_TH_PRODUCER ? 32 : 0);
constexpr int TH_ISSUER = SB_TH_PRODUCER ? TH_COLS : 0;      // thread that issues the bulk copies
template <typename T, bool SUBST, bool FROM_VEC, int KIND, bool TAB>
__global__ void __launch_bounds__(TH_THREADS)
thomas_sweep(ThomasTab tb, int strip_first, const T* __restrict__ in, const T* __restrict__ V,
             const double* __restrict__ gvec, const float* __restrict__ gvecf,
             const double* __restrict__ bsig, T* __restrict__ out) {
  constexpr int NS = ThStages<T, KIND, TAB>::v;
  constexpr int RT = ThRows<T, KIND, TAB>::v;
  constexpr bool PLAIN = sizeof(T) == 4 && !TAB;     // fp32 arithmetic, float coefficient rows
  constexpr bool RECOMP = PLAIN && KIND == 2;        // E is recomputed from checkpoints, not loaded
  constexpr bool COMBINE = KIND == 2 && !RECOMP;     // a second operand tile (E) travels with the first
  constexpr bool GVEC = FROM_VEC || RECOMP;          // rows of the shared right-hand side g are staged
  static_assert(SUBST || KIND == 0, "KIND applies to substitution sweeps");
  constexpr bool LOAD = !FROM_VEC;
  static_assert(TH_COLS == SP_W, "one CTA per 64-column strip");
  static_assert(RT % TH_RT == 0, "staged tile = whole register blocks");
  constexpr size_t TILE_B = (size_t)RT * TH_COLS * sizeof(T);
  constexpr size_t CT_B = TAB ? (size_t)RT * TH_COLS * sizeof(double) : 0;
  extern __shared__ __align__(128) unsigned char th_smem[];
  double (*tileC)[RT][TH_COLS] = reinterpret_cast<double (*)[RT][TH_COLS]>(th_smem);
  T (*tileA)[RT][TH_COLS] = reinterpret_cast<T (*)[RT][TH_COLS]>(th_smem + NS * CT_B);
  T (*tileV)[RT][TH_COLS] = reinterpret_cast<T (*)[RT][TH_COLS]>(th_smem + NS * CT_B + NS * TILE_B);
  // fp64 side tiles [RT][KB] for the indefinite columns (low-k class only; KB even)
  double* tileD = reinterpret_cast<double*>(th_smem + NS * CT_B + NS * TILE_B * (COMBINE ? 2 : 1));
  const size_t SIDE_B = TAB ? (size_t)RT * tb.KB * sizeof(double) : 0;
  double* tileD1 = tileD + (COMBINE ? NS * SIDE_B / sizeof(double) : 0);      // side tiles of E (KIND 2)
  unsigned long long* full = reinterpret_cast<unsigned long long*>(
      th_smem + NS * CT_B + NS * TILE_B * (COMBINE ? 2 : 1) + NS * SIDE_B * (COMBINE ? 2 : 1));
  const bool worker = threadIdx.x < TH_COLS;            // owns a column (the producer warp does not)
  const bool issuer = threadIdx.x == TH_ISSUER;
  const int tid = worker ? threadIdx.x : 0;             // column of the thread inside the strip
  const int strip = strip_first + blockIdx.x;
  const int c = strip * TH_COLS + tid;
  const int plane = blockIdx.y, m = plane % tb.nl;
  const int half = blockIdx.z & 1, seg = blockIdx.z >> 1;
  const bool act = worker && c < tb.ncols;
  const int ny = tb.ny;
  const size_t strip0 = (size_t)plane * ny * tb.np + (size_t)strip * ny * SP_W;   // strip base
  const double cfix = tb.cinf[(size_t)m * tb.np + c];
  const int Js = tb.Jstrip[m * tb.nstrip + strip];
  const double* tabS = tb.ctabB + tb.tabOff[m * tb.nstrip + strip];
  const bool bad = TAB && act && c < tb.kbad[m];
  const bool strip_bad = TAB && strip * TH_COLS < tb.kbad[m];     // this strip holds indefinite columns
  // side buffer / meeting values written by this elimination, or read by this substitution
  double* sideG = (FROM_VEC ? tb.dbad1 : tb.dbad) + ((size_t)plane * ny) * tb.KB;   // [j][KB] of this plane
  const double* sideG1 = tb.dbad1 + ((size_t)plane * ny) * tb.KB;
  double* meetW = FROM_VEC ? tb.meet1 : tb.meet;
  const double bs = (KIND == 2 && act) ? bsig[c] : 0.0;
  const float cfix_f = (float)cfix, kfix_f = (float)(cfix * tb.dy2), dy2_f = (float)tb.dy2, bs_f = (float)bs;
  float carry_f = 0.f;
  // FROM_VEC: the shared right-hand side g[j] of the tile after next is fetched (one element per
  // thread, coalesced) while the current tile runs, and read back as a shared-memory broadcast
  using GT = typename std::conditional<PLAIN, float, double>::type;
  __shared__ GT gbuf[2][GVEC ? 2 * RT : 1];      // [buffer][column parity][row of the tile]
  const GT* gsrc = nullptr;
  if (GVEC) {
    if constexpr (PLAIN) gsrc = gvecf + (size_t)plane * 2 * ny; else gsrc = gvec + (size_t)plane * 2 * ny;
  }
  const int gpar = GVEC ? (c & 1) * RT : 0;
  const int m1 = ny / 2;
  const int cnt = half == 0 ? m1 : ny - m1;
  int j0, dj;
  if (!SUBST) { j0 = half == 0 ? 0 : ny - 1; dj = half == 0 ? 1 : -1; }
  else        { j0 = half == 0 ? m1 - 1 : m1; dj = half == 0 ? -1 : 1; }
  // this CTA's segment [sa, sb) of the half's sequence (the whole half when nseg = 1)
  const int sa0 = min(cnt, seg * tb.seg_len), sb = tb.nseg > 1 ? min(cnt, sa0 + tb.seg_len) : cnt;
  const bool probe = tb.pass == 1;
  // The probe pass of a PLAIN fp32 strip only needs the tail of its segment: the recurrence forgets
  // its start like |c|^rows (|c| <= ~0.9 on these strips), so the end value of a run from zero over
  // the last `vwarm` rows (|c_max|^rows < 1e-13, the table thomas_vec_ckpt uses) is the end value of
  // the whole segment to far below fp32 rounding - and the pass reads that much less.
  int sa = sa0;
  if (PLAIN && probe && tb.nseg > 1) {
    const int wrm = tb.vwarm[m * tb.nstrip + strip];
    if (sb - sa0 > wrm) sa = sa0 + ((sb - wrm - sa0) / RT) * RT;
  }
  const int ntile = (sb - sa + RT - 1) / RT;
  // tile t covers sequence numbers s0..s0+nr-1, i.e. memory rows jlo..jlo+nr-1 (ascending), and
  // table indices ilo..ilo+nr-1 (i = s for elimination, cnt-1-s for substitution)
  auto tile_nr = [&](int t) { return min(RT, sb - (sa + t * RT)); };
  auto tile_jlo = [&](int t) {
    const int s0 = sa + t * RT, nr = min(RT, sb - s0);
    return dj > 0 ? j0 + s0 : j0 - (s0 + nr - 1);
  };
  auto tile_ilo = [&](int t) {
    const int s0 = sa + t * RT, nr = min(RT, sb - s0);
    return SUBST ? cnt - 1 - (s0 + nr - 1) : s0;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  auto load_tile = [&](int t) {       // executed by thread 0 only
    if (t >= ntile) return;
    const int st = t % NS;
    const int nr = tile_nr(t), ilo = tile_ilo(t);
    const int ncoef = TAB ? max(0, min(ilo + nr, Js) - ilo) : 0;      // tabulated rows of this tile
    const unsigned bytes = (unsigned)(nr * TH_COLS * sizeof(T));
    const unsigned cbytes = (unsigned)(ncoef * TH_COLS * sizeof(double));
    const unsigned dbytes = (TAB && SUBST && strip_bad) ? (unsigned)(nr * tb.KB * sizeof(double)) : 0u;
    const unsigned total = (LOAD ? bytes * (COMBINE ? 2u : 1u) : 0u) + cbytes + dbytes * (COMBINE ? 2u : 1u);
    if (total == 0) return;
    const size_t off = strip0 + (size_t)tile_jlo(t) * SP_W;
    mbar_arrive_expect_tx(&full[st], total);
    if (dbytes) bulk_g2s(tileD + (size_t)st * RT * tb.KB, sideG + (size_t)tile_jlo(t) * tb.KB, dbytes, &full[st]);
    if (dbytes && COMBINE)
      bulk_g2s(tileD1 + (size_t)st * RT * tb.KB, sideG1 + (size_t)tile_jlo(t) * tb.KB, dbytes, &full[st]);
    if (LOAD) bulk_g2s(&tileA[st][0][0], in + off, bytes, &full[st]);
    if (COMBINE) bulk_g2s(&tileV[st][0][0], V + off, bytes, &full[st]);
    if (cbytes) bulk_g2s(&tileC[st][0][0], tabS + (size_t)ilo * TH_COLS, cbytes, &full[st]);
  };
  auto tile_has_load = [&](int t) {
    return LOAD || (TAB && tile_ilo(t) < Js) || (TAB && SUBST && strip_bad);
  };
  if (issuer)
    for (int t = 0; t < NS - 1; ++t) load_tile(t);

  double carry = 0.0;
  if (SUBST && cnt > 0 && m1 > 0) {
    // meeting point: x_{m1-1} + ca x_{m1} = d_{m1-1};  x_{m1} + cb x_{m1-1} = e_{m1}
    const double ca = tb.meetc[((size_t)m * 2 + 0) * tb.np + c], cb = tb.meetc[((size_t)m * 2 + 1) * tb.np + c];
    double dm = tb.meet[((size_t)plane * 2 + 0) * tb.np + c];
    double em = tb.meet[((size_t)plane * 2 + 1) * tb.np + c];
    if (KIND == 2) {      // eliminated D - bs E at the meeting rows
      dm = fma(-bs, tb.meet1[((size_t)plane * 2 + 0) * tb.np + c], dm);
      em = fma(-bs, tb.meet1[((size_t)plane * 2 + 1) * tb.np + c], em);
    }
    const double den = 1.0 / (1.0 - ca * cb);
    const double xa = (dm - ca * em) * den, xb = (em - cb * dm) * den;
    carry = half == 0 ? xb : xa;     // the neighbour's value across the meeting point
    carry_f = (float)carry;
  }
  if (tb.nseg > 1) {
    if (probe) {
      carry = 0.0;                    // pass 1: every segment from a zero carry
    } else {
      const double* eb = tb.segbuf + ((size_t)(plane * 2 + half) * tb.nseg) * tb.np + c;
      const double* pb = tb.segprod + ((size_t)(((SUBST ? 2 : 0) + half) * tb.nl + m) * tb.nseg) * tb.np + c;
      for (int k = 0; k < seg; ++k) carry = fma(pb[(size_t)k * tb.np], carry, eb[(size_t)k * tb.np]);
    }
    carry_f = (float)carry;
  }
  // checkpoint of the tile's E block (RECOMP): block index by the elimination rows it covers
  const float* ckS = nullptr;
  auto tile_ck = [&](int t) -> float {
    const int s1 = sa + t * RT + tile_nr(t);                  // rows [cnt - s1, ...) in elimination order
    return (s1 >= cnt) ? 0.f : ckS[(size_t)(s1 / RT - 1) * TH_COLS + tid];
  };
  if (RECOMP) ckS = tb.ckpt + ((size_t)(plane * tb.nstrip + strip) * 2 + (half == 0 ? 0 : 1)) * tb.nck * TH_COLS;
  float ck_cur = (RECOMP && worker && ntile > 0) ? tile_ck(0) : 0.f;
  if (GVEC && ntile > 0) {
    if (worker && tid < tile_nr(0)) {
      gbuf[0][tid] = gsrc[tile_jlo(0) + tid];
      gbuf[0][RT + tid] = gsrc[ny + tile_jlo(0) + tid];
    }
    __syncthreads();
  }
  T wdot[KIND == 1 ? 16 : 1];
  if (KIND == 1) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      wdot[i] = reinterpret_cast<const T*>(tb.sig2n)[strip * TH_COLS + (tid >> 4) * 16 + ((i + (tid & 15)) & 15)];
  }
  unsigned phase_bits = 0;            // per-stage phase parity (stages may be skipped by FROM_VEC tiles)
#ifdef SB_TH_PHASES
  long long ph_wait = 0, ph_comp = 0, ph_sync = 0;
  const long long ph_begin = clock64();
#endif
#ifdef SB_TH_DEBUG
  unsigned long long tg0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tg0));
#endif
#pragma unroll 1
  for (int t = 0; t < ntile; ++t) {
    const int st = t % NS;
    const int nrt = tile_nr(t);
    const int ilot = tile_ilo(t);
    GT gnext = 0, gnext1 = 0;
    if (GVEC && worker && t + 1 < ntile && tid < tile_nr(t + 1)) {
      gnext = gsrc[tile_jlo(t + 1) + tid];
      gnext1 = gsrc[ny + tile_jlo(t + 1) + tid];
    }
    const float ck_next = (RECOMP && worker && t + 1 < ntile) ? tile_ck(t + 1) : 0.f;
    const GT* gvt = gbuf[t & 1] + gpar - tile_jlo(t);      // indexed by the memory row
#ifdef SB_TH_PHASES
    long long ph0 = clock64();
#endif
    if (worker && tile_has_load(t)) {
      mbar_wait(&full[st], (phase_bits >> st) & 1u);
      phase_bits ^= 1u << st;
    }
#ifdef SB_TH_PHASES
    long long ph1 = clock64();
    ph_wait += ph1 - ph0;
#endif
#pragma unroll 1
    for (int b0 = 0; worker && b0 < nrt; b0 += TH_RT) {       // register blocks, in sequence order
      const int nr = min(TH_RT, nrt - b0);
      const int s0 = sa + t * RT + b0;
      const int ilo = SUBST ? cnt - 1 - (s0 + nr - 1) : s0;
      const int row0 = dj > 0 ? b0 : nrt - b0 - nr;   // first memory row of the block inside the tile
      T (*A)[TH_COLS] = tileA[st] + row0;
      T (*Vt)[TH_COLS] = tileV[st] + row0;
      const int jb = j0 + dj * s0;
      // mode: 0 = constant coefficient, 1 = fully tabulated, 2 = generic
      const int mode = (nr < TH_RT || (ilo < Js && ilo + nr > Js)) ? 2 : (ilo >= Js ? 0 : 1);
      if constexpr (PLAIN) {
        // float coefficient rows, read straight from L2 (only the first Js rows of a half)
        const float (*Ct)[TH_COLS] = reinterpret_cast<const float (*)[TH_COLS]>(tabS) + ilo;
#define SB_TILE(UPV, MODEV)                                                                        \
        thomas_tile_f32<SUBST, FROM_VEC, KIND, UPV, MODEV>(                                        \
            reinterpret_cast<float (*)[TH_COLS]>(A), reinterpret_cast<const float (*)[TH_COLS]>(Vt), Ct, (const float*)gvt, \
            tid, nr, s0, ilo, Js, cnt, jb, act, cfix_f, kfix_f, dy2_f, bs_f, ck_cur, carry_f)
        if (dj > 0) {
          if (mode == 0) SB_TILE(true, 0); else if (mode == 1) SB_TILE(true, 1); else SB_TILE(true, 2);
        } else {
          if (mode == 0) SB_TILE(false, 0); else if (mode == 1) SB_TILE(false, 1); else SB_TILE(false, 2);
        }
#undef SB_TILE
      } else {
        // coefficient rows of this block: staged in shared memory (TAB) or read straight from L2
        const double (*Ct)[TH_COLS] = TAB ? (const double (*)[TH_COLS])(tileC[st] + (ilo - ilot))
                                          : (const double (*)[TH_COLS])(tabS + (size_t)ilo * TH_COLS);
        double* Dt = tileD + ((size_t)st * RT + row0) * tb.KB + (bad ? c : 0);
        const double* Dt1 = tileD1 + ((size_t)st * RT + row0) * tb.KB + (bad ? c : 0);
#define SB_TILE(UPV, MODEV)                                                                        \
        thomas_tile<T, SUBST, FROM_VEC, KIND, UPV, MODEV>(A, Vt, Ct, (const double*)gvt, Dt, Dt1, tb.KB, tid, nr, s0, ilo, Js, \
                                                             cnt, jb, act, bad, cfix, tb.dy2, bs, carry)
        if (dj > 0) {
          if (mode == 0) SB_TILE(true, 0); else if (mode == 1) SB_TILE(true, 1); else SB_TILE(true, 2);
        } else {
          if (mode == 0) SB_TILE(false, 0); else if (mode == 1) SB_TILE(false, 1); else SB_TILE(false, 2);
        }
#undef SB_TILE
      }
    }
#ifdef SB_TH_PHASES
    long long ph2 = clock64();
    ph_comp += ph2 - ph1;
#endif
    if (GVEC && worker && t + 1 < ntile && tid < tile_nr(t + 1)) {
      gbuf[(t + 1) & 1][tid] = gnext;
      gbuf[(t + 1) & 1][RT + tid] = gnext1;
    }
    ck_cur = ck_next;
    if (worker && !SUBST && !probe && t == ntile - 1 && sb == cnt)
      meetW[((size_t)plane * 2 + half) * tb.np + c] = PLAIN ? (double)carry_f : carry;
    // finished tile -> global (the bulk store reads shared memory through the async proxy)
    fence_async_smem();
    __syncthreads();
    if (KIND == 1 && worker && !probe) {
      // border sums of the finished tile: thread (rr, qd) adds 16 columns of row rr with a rotated
      // column order (conflict-free), the two quarters of a warp combine by shuffle
      const int rr = tid & 15, qd = tid >> 4;
      const size_t pstride = (size_t)2 * tb.nstrip * ny;      // even-column block -> odd-column block
      T* part = reinterpret_cast<T*>(tb.part) +
                ((size_t)plane * 4 * tb.nstrip + strip * 2 + (tid >> 5)) * ny + tile_jlo(t);
#pragma unroll 1
      for (int rb = 0; rb < nrt; rb += 16) {
        const int row = rb + rr;
        // column (i + rr) & 15 has the parity of i + rr: even i and odd i accumulate separately
        T acc0 = 0, acc1 = 0;
        if (row < nrt) {
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            acc0 = fma(wdot[i], tileA[st][row][qd * 16 + ((i + rr) & 15)], acc0);
            acc1 = fma(wdot[i + 1], tileA[st][row][qd * 16 + ((i + 1 + rr) & 15)], acc1);
          }
        }
        T ev = (rr & 1) ? acc1 : acc0, od = (rr & 1) ? acc0 : acc1;
        ev += __shfl_xor_sync(0xffffffffu, ev, 16);
        od += __shfl_xor_sync(0xffffffffu, od, 16);
        if ((tid & 16) == 0 && row < nrt) { part[row] = ev; part[pstride + row] = od; }
      }
    }
#ifdef SB_TH_PHASES
    ph_sync += clock64() - ph2;
#endif
    if (issuer) {
      if (KIND == 2 && !probe && tb.pn > 0) {
        // rows [jlo, jlo + nrt) of the tile, cut at the row ownership boundaries
        const int jlo = tile_jlo(t);
        for (int r = 0; r < tb.pn; ++r) {
          const int a0 = max(jlo, tb.prow[r]), a1 = min(jlo + nrt, tb.prow[r + 1]);
          if (a1 <= a0) continue;
          const int rr = tb.prow[r + 1] - tb.prow[r];
          T* d = reinterpret_cast<T*>(tb.ppeer[r]) + (((size_t)plane * tb.nstrip + strip) * rr + (a0 - tb.prow[r])) * SP_W;
          bulk_s2g(d, &tileA[st][a0 - jlo][0], (unsigned)((a1 - a0) * TH_COLS * sizeof(T)));
        }
      } else if (KIND != 1 && !probe)
        bulk_s2g(out + strip0 + (size_t)tile_jlo(t) * SP_W, &tileA[st][0][0], (unsigned)(nrt * TH_COLS * sizeof(T)));
      if (TAB && !SUBST && strip_bad && !probe)
        bulk_s2g(sideG + (size_t)tile_jlo(t) * tb.KB, tileD + (size_t)st * RT * tb.KB,
                 (unsigned)(nrt * tb.KB * sizeof(double)));
      bulk_commit();
      // stage (t-1)%NS is free once the store of tile t-1 has finished reading shared memory
      bulk_wait_read<1>();
      load_tile(t + NS - 1);
    }
  }
  if (issuer) bulk_wait_read<0>();
  if (probe && worker)
    tb.segbuf[((size_t)(plane * 2 + half) * tb.nseg + seg) * tb.np + c] = PLAIN ? (double)carry_f : carry;
#ifdef SB_TH_PHASES
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    const unsigned slot = atomicAdd(&g_dbg_n, 1u);
    if (slot < 4096) {
      g_dbg[slot * 4 + 0] = ((SUBST ? 1000 : 0) + (FROM_VEC ? 100 : 0) + KIND * 10 + (TAB ? 1 : 0)) * 10 + half +
                            ((unsigned long long)(clock64() - ph_begin) << 32);
      g_dbg[slot * 4 + 1] = ph_wait; g_dbg[slot * 4 + 2] = ph_comp; g_dbg[slot * 4 + 3] = ph_sync;
    }
  }
#endif
#ifdef SB_TH_DEBUG
  if (threadIdx.x == 0 && (TAB || (blockIdx.x % 41 == 0)) && plane < tb.nl) {
    unsigned long long tg1; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tg1));
    const unsigned slot = atomicAdd(&g_dbg_n, 1u);
    if (slot < 4096) {
      g_dbg[slot * 4 + 0] = (SUBST ? 1000 : 0) + (FROM_VEC ? 100 : 0) + KIND * 10 + (TAB ? 1 : 0);
      g_dbg[slot * 4 + 1] = strip * 100 + plane * 10 + half;
      g_dbg[slot * 4 + 2] = tg0; g_dbg[slot * 4 + 3] = tg1;
    }
  }
#endif
}

// Elimination of the border right-hand side of the second solve on the PLAIN fp32 strips, WITHOUT
// storing it: every column runs d_i = (dy^2 g_i - d_{i-1}) c_i over its half (g = gvecf[c & 1], staged
// in shared memory) and keeps d_{i-1} whenever (cnt - i) is a multiple of TH_RT - the value the
// KIND 2 substitution needs to recompute its block of E - plus the last value for the meeting-point
// solve.  No array transfer at all (the stored version wrote and re-read one).
//
// The half is cut into VC_NSEG segments run by different CTAs.  The recurrence forgets its start
// like |c|^rows and the PLAIN strips have |c| <= ~0.9, so a segment starts from zero `warm` rows
// early (warm = rows until |c_max|^rows < 1e-13, per strip, tabulated on the host; a segment whose
// warm-up reaches the start of the half starts there, exactly): what is left of the wrong start is far
// below fp32 rounding, and the serial chain is VC_NSEG times shorter (it sits on the critical path
// of the slab-distributed model: 0.066 ms per evaluation at 8192^2 on 8 GPUs, unsegmented).
constexpr int VC_NSEG = 8;

__global__ void __launch_bounds__(TH_COLS)
thomas_vec_ckpt(ThomasTab tb, int strip_first, const float* __restrict__ gvecf, int nseg) {
  extern __shared__ float gsh[];                 // [2][rows of this segment incl. warm-up], elimination order
  const int tid = threadIdx.x, strip = strip_first + blockIdx.x, c = strip * TH_COLS + tid;
  const int plane = blockIdx.y, m = plane % tb.nl, half = blockIdx.z & 1, seg = blockIdx.z >> 1, ny = tb.ny;
  const int m1 = ny / 2, cnt = half == 0 ? m1 : ny - m1;
  const int j0 = half == 0 ? 0 : ny - 1, dj = half == 0 ? 1 : -1;
  // segment boundaries sit on the checkpoint grid: (cnt - sa) is a multiple of TH_RT; segment 0 also
  // takes the ragged rows [0, cnt mod TH_RT)
  const int r0 = cnt % TH_RT, nblk = (cnt - r0) / TH_RT;
  const int sa = seg == 0 ? 0 : r0 + TH_RT * (int)((long long)seg * nblk / nseg);
  const int sb = r0 + TH_RT * (int)((long long)(seg + 1) * nblk / nseg);
  if (sb <= sa) return;
  const int warm = tb.vwarm[m * tb.nstrip + strip];
  const int i0 = (seg == 0 || sa - warm <= r0) ? 0 : sa - warm;      // warm is a multiple of TH_RT
  const int nrow = sb - i0;
  const float* g = gvecf + (size_t)plane * 2 * ny;
  for (int i = tid; i < nrow; i += TH_COLS) {
    gsh[i] = g[j0 + dj * (i0 + i)];
    gsh[nrow + i] = g[ny + j0 + dj * (i0 + i)];
  }
  __syncthreads();
  const float* gs = gsh + (c & 1) * nrow - i0;     // indexed by the elimination row
  const double cfix = tb.cinf[(size_t)m * tb.np + c];
  const float cfix_f = (float)cfix, kfix_f = (float)(cfix * tb.dy2), dy2_f = (float)tb.dy2;
  const float kinv_f = kfix_f != 0.f ? 1.f / kfix_f : 0.f;      // (padding columns have c = 0)
  const int Js = tb.Jstrip[m * tb.nstrip + strip];
  const float (*Ct)[TH_COLS] = reinterpret_cast<const float (*)[TH_COLS]>(tb.ctabB + tb.tabOff[m * tb.nstrip + strip]);
  float* ck = tb.ckpt + ((size_t)(plane * tb.nstrip + strip) * 2 + half) * tb.nck * TH_COLS + tid;
  float carry = 0.f;
  int i = i0;
  if (i0 == 0)
    for (; i < r0; ++i) {      // the ragged block first
      const float cj = i < Js ? Ct[i][tid] : cfix_f;
      carry = fmaf(-cj, carry, (i < Js ? cj * dy2_f : kfix_f) * gs[i]);
    }
  for (; i < sb; i += TH_RT) {
    if (i >= sa) ck[(size_t)((cnt - i) / TH_RT - 1) * TH_COLS] = carry;      // (not during the warm-up)
    float gg[TH_RT];
#pragma unroll
    for (int r = 0; r < TH_RT; ++r) gg[r] = gs[i + r];
    if (i >= Js) {
      // constant coefficient: u = carry / k runs u' = g - c u, one FMA per row on the chain
      float u = carry * kinv_f;
#pragma unroll
      for (int r = 0; r < TH_RT; ++r) u = fmaf(-cfix_f, u, gg[r]);
      carry = u * kfix_f;
    } else {
#pragma unroll
      for (int r = 0; r < TH_RT; ++r) {
        const bool tab = i + r < Js;
        const float cj = tab ? Ct[i + r][tid] : cfix_f;
        carry = fmaf(-cj, carry, (tab ? cj * dy2_f : kfix_f) * gg[r]);
      }
    }
  }
  if (sb == cnt && c < tb.np) tb.meet1[((size_t)plane * 2 + half) * tb.np + c] = (double)carry;
}

// Right-hand sides of the border (Schur) system.  rvec / ghat are interleaved so that one 128-bit
// load fetches two columns of a row: element (batch b, row j, vector v = 3 * layer + i) lives at
// (b * ny + j) * NVP + v, NVP = border_nvp(nl).
//   r0 = f_0 - b v(1),  r1 = f_n - b v(n-1),  r2 = f_{n+1}
// with v(1) = E + O and v(n-1) = E - O, E / O = sum over the even / odd columns c (odd / even
// x-wavenumbers k = c + 1) of sig2n[c] x[j][c].  The sums arrive as 2 * nstrip per-warp partials per
// parity from the KIND 1 substitution sweeps (fixed order: deterministic); f_n sits in the border
// slot of S, f_0 and f_{n+1} in bext.
__host__ __device__ inline int border_nvp(int nl) { return (3 * nl + 1) & ~1; }

// Thread (jx, py) of a CTA adds the partials p = py, py + 8, ... of row j0 + jx (loads coalesced over
// jx), the eight py combine through shared memory in a fixed order.  (One thread per row walked
// all 2 * 256 partials alone: 0.09 ms at 8192^2 on 99 CTAs.)
template <typename T>
__global__ void __launch_bounds__(256)
border_reduce(const T* __restrict__ S, const T* __restrict__ bext, const T* __restrict__ part,
              int ny, int np, int ncols, int npart, int nl, double b, double* __restrict__ r) {
  __shared__ double red[2][8][32];
  const int jx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jx, plane = blockIdx.y;
  double ev = 0, od = 0;
  if (j < ny) {
    const T* pe = part + (size_t)plane * 2 * npart * ny + j;
    const T* po = pe + (size_t)npart * ny;
#pragma unroll 8
    for (int p = py; p < npart; p += 8) { ev += (double)pe[(size_t)p * ny]; od += (double)po[(size_t)p * ny]; }
  }
  red[0][py][jx] = ev; red[1][py][jx] = od;
  __syncthreads();
  if (py != 0 || j >= ny) return;
  ev = 0; od = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) { ev += red[0][q][jx]; od += red[1][q][jx]; }
  const int bm = plane / nl, l = plane - bm * nl;
  double* rp = r + ((size_t)bm * ny + j) * border_nvp(nl) + 3 * l;
  rp[0] = (double)bext[((size_t)plane * 3 + 0) * ny + j] - b * (ev + od);
  rp[1] = (double)bext[((size_t)plane * 3 + 2) * ny + j] - b * (ev - od);
  rp[2] = (double)bext[((size_t)plane * 3 + 1) * ny + j];
}

// Slab-distributed model: the per-warp partials of the strips [s0, s1) a rank owns, summed per
// (plane, parity, row) into the rank's slot of pvec (fixed order), so that 2 * planes vectors leave
// the rank instead of 4 * (s1 - s0) * planes; border_reduce_ranks then adds the ranks' vectors in rank
// order.  (All partials to every rank: 44 MB per rank and evaluation at 8192^2 on 8 GPUs.)
template <typename T>
__global__ void __launch_bounds__(256)
border_presum(const T* __restrict__ part, int ny, int npart, int p0, int p1, double* __restrict__ pv) {
  __shared__ double red[8][32];
  const int jx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jx, pp = blockIdx.y;            // pp = plane * 2 + parity
  double acc = 0;
  if (j < ny) {
    const T* src = part + (size_t)pp * npart * ny + j;
    for (int p = p0 + py; p < p1; p += 8) acc += (double)src[(size_t)p * ny];
  }
  red[py][jx] = acc;
  __syncthreads();
  if (py != 0 || j >= ny) return;
  acc = 0;
#pragma unroll
  for (int q = 0; q < 8; ++q) acc += red[q][jx];
  pv[(size_t)pp * ny + j] = acc;
}

template <typename T>
__global__ void __launch_bounds__(256)
border_reduce_ranks(const T* __restrict__ bext, const double* __restrict__ pvec, int nranks, int planes,
                    int ny, int nl, double b, double* __restrict__ r) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, plane = blockIdx.y;
  if (j >= ny) return;
  double ev = 0, od = 0;
  for (int q = 0; q < nranks; ++q) {
    const double* pv = pvec + ((size_t)q * planes + plane) * 2 * ny + j;
    ev += pv[0]; od += pv[ny];
  }
  const int bm = plane / nl, l = plane - bm * nl;
  double* rp = r + ((size_t)bm * ny + j) * border_nvp(nl) + 3 * l;
  rp[0] = (double)bext[((size_t)plane * 3 + 0) * ny + j] - b * (ev + od);
  rp[1] = (double)bext[((size_t)plane * 3 + 2) * ny + j] - b * (ev - od);
  rp[2] = (double)bext[((size_t)plane * 3 + 1) * ny + j];
}

// Epilogue of the border transforms for output row a (1-based) of one layer's three columns.
// STAGE_A multiplies by the host-inverted 3 x 3 Schur block of y-mode a; stage B scales by 2/N and
// writes the results where the sweeps / inverse transform read them: gvec = (g0 + g1, g0 - g1), g1
// in the border slot of S, g0 and g2 in bext.
template <typename T, bool STAGE_A>
__device__ __forceinline__ void border_store(double t0, double t1, double t2, int a, int bm, int l, int nl,
                                             const double* __restrict__ minv, int ny, int np, int n,
                                             double* __restrict__ out, double* __restrict__ gvec,
                                             float* __restrict__ gvecf, T* __restrict__ S,
                                             T* __restrict__ bext) {
  const int plane = bm * nl + l;
  if (STAGE_A) {
    const double* M = minv + ((size_t)l * ny + (a - 1)) * 9;
    double* o = out + ((size_t)bm * ny + (a - 1)) * border_nvp(nl) + 3 * l;
    o[0] = M[0] * t0 + M[1] * t1 + M[2] * t2;
    o[1] = M[3] * t0 + M[4] * t1 + M[5] * t2;
    o[2] = M[6] * t0 + M[7] * t1 + M[8] * t2;
  } else {
    const double sc = 2.0 / (ny + 1);
    t0 *= sc; t1 *= sc; t2 *= sc;
    double* gv = gvec + (size_t)plane * 2 * ny + (a - 1);
    gv[0] = t0 + t1; gv[ny] = t0 - t1;
    if (sizeof(T) == 4) {
      float* gf = gvecf + (size_t)plane * 2 * ny + (a - 1);
      gf[0] = (float)(t0 + t1); gf[ny] = (float)(t0 - t1);
    }
    S[(size_t)plane * ny * np + sp_off(ny, a - 1, n - 1)] = (T)t1;
    bext[((size_t)plane * 3 + 0) * ny + (a - 1)] = (T)t0;
    bext[((size_t)plane * 3 + 1) * ny + (a - 1)] = (T)t2;
  }
}

// Border Schur solve = dense DST-I in y of the 3 nl border columns of a member (length ny, N = ny + 1),
// brute force in fp64: out_v[a] = sum_{t=1..ny} sin(pi a t / N) in_v[t].
// A CTA computes GS_OUT = 32 outputs a OF ONE PARITY (a, a + 2, ...) for ALL NV columns: lane = output,
// warp = one of GS_SL interleaved slices of the terms t, so every load of in[t][.] is a warp-wide
// broadcast and the rotation that advances sin/cos(pi a t / N) by the fixed angle pi a GS_SL / N
// (4 FMAs; start and step values come exactly from the table) is shared by the NV columns.  The
// input is folded by the t <-> N - t symmetry, sin(pi a (N - t) / N) = -(-1)^a sin(pi a t / N): half
// the terms, and because the sign is the same for the whole CTA the fold x[t] -+ x[N-t] is done once
// per staged chunk in shared memory (13 fp64 operations per output and term instead of 22 when every
// lane folded for itself).  The slices are combined through shared memory in a fixed order.  (The first version ran one warp per
// (output, layer): every warp re-read its three input columns from L2, 4.7 GB per launch at 8192^2.)
constexpr int GS_OUT = 32, GS_SL = 8;
constexpr int GS_U = 8;                  // terms per slice and chunk
constexpr int GS_CH = GS_SL * GS_U;      // rows of the input per staged chunk
constexpr int GS_NS = 3;                 // ring depth
constexpr int GS_ZMAX = 8;               // at most this many term-split CTAs per output block

template <typename T, int NL, bool STAGE_A>
__global__ void __launch_bounds__(GS_OUT * GS_SL)
border_dst(const double* __restrict__ in, const double* __restrict__ sintab,
           const double* __restrict__ minv, int ny, int np, int n,
           double* __restrict__ out, double* __restrict__ gvec, float* __restrict__ gvecf,
           T* __restrict__ S, T* __restrict__ bext, int a_first, int a_count,
           double* __restrict__ zpart, unsigned* __restrict__ zcount) {
  constexpr int NV = 3 * NL, NVP = (NV + 1) & ~1;
  // gridDim.z > 1: the terms are split over gridDim.z CTAs per output block (few output blocks - a
  // rank of the slab model computes a slice of the outputs - would leave most SMs idle and every
  // CTA with the full-length chain); partial sums meet in `zpart`, the CTA that arrives last adds
  // them in z order (deterministic) and runs the epilogue.
  // Input rows t-1 (lo) and N-t-1 (hi) of a chunk of terms are two contiguous blocks of the
  // interleaved array: one cp.async.bulk each into a GS_NS-stage ring, completion on an mbarrier
  // (first version: per-term global loads, every one an L2 round trip with no memory-level
  // parallelism - 0.39 ms per launch at 8192^2, 10 % of the DFMA rate).
  constexpr size_t TILE_B = sizeof(double) * GS_NS * 2 * GS_CH * NVP, RED_B = sizeof(double) * GS_SL * NV * GS_OUT;
  __shared__ __align__(128) unsigned char gs_smem[TILE_B > RED_B ? TILE_B : RED_B];
  __shared__ unsigned long long full[GS_NS];
  double (*tile)[2][GS_CH][NVP] = reinterpret_cast<double (*)[2][GS_CH][NVP]>(gs_smem);
  double (*red)[NV][GS_OUT] = reinterpret_cast<double (*)[NV][GS_OUT]>(gs_smem);   // after the last chunk
  const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5, bm = blockIdx.y;
  const int a = a_first + ((int)blockIdx.x >> 1) * (2 * GS_OUT) + 2 * lane + ((int)blockIdx.x & 1) + 1;      // 1-based output row
  const bool valid = a <= a_first + a_count;
  const unsigned N = ny + 1, N2 = 2 * N;
  const double* costab = sintab + N2;
  const unsigned aa = valid ? (unsigned)a : 1u;
  const double* x = in + (size_t)bm * ny * NVP;
  const int half = (N - 1) / 2;                       // terms t = 1..half (folded)
  const int nchunk_all = (half + GS_CH - 1) / GS_CH;
  const int nz = gridDim.z, z = blockIdx.z;
  const int ch_first = (int)((long long)z * nchunk_all / nz), nchunk = (int)((long long)(z + 1) * nchunk_all / nz);
  const unsigned k0 = (unsigned)(((unsigned long long)aa * (unsigned)(ch_first * GS_CH + sl + 1)) % N2);
  const unsigned kd = (aa * (unsigned)GS_SL) % N2;
  double s = sintab[k0], c = costab[k0];
  const double ds = sintab[kd], dc = costab[kd];
  const double sgn = ((a_first + ((int)blockIdx.x & 1) + 1) & 1) ? 1.0 : -1.0;      // the CTA's parity (also of its invalid lanes)
  double acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = 0.0;
  auto chunk_rows = [&](int ch) { return min(GS_CH, half - ch * GS_CH); };
  auto load_chunk = [&](int ch) {                     // thread 0 only
    if (ch >= nchunk) return;
    const int st = (ch - ch_first) % GS_NS, nr = chunk_rows(ch), t0 = ch * GS_CH + 1;
    const unsigned bytes = (unsigned)(nr * NVP * sizeof(double));
    mbar_arrive_expect_tx(&full[st], 2 * bytes);
    bulk_g2s(&tile[st][0][0][0], x + (size_t)(t0 - 1) * NVP, bytes, &full[st]);
    bulk_g2s(&tile[st][1][0][0], x + (size_t)(N - t0 - nr) * NVP, bytes, &full[st]);   // rows N-t-1, ascending
  };
  if (threadIdx.x == 0) {
    for (int q = 0; q < GS_NS; ++q) mbar_init(&full[q], 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    for (int ch = ch_first; ch < ch_first + GS_NS - 1; ++ch) load_chunk(ch);
  }
  __syncthreads();
  for (int ch = ch_first; ch < nchunk; ++ch) {
    const int st = (ch - ch_first) % GS_NS, nr = chunk_rows(ch);
    mbar_wait(&full[st], (unsigned)(((ch - ch_first) / GS_NS) & 1));
    // fold in place: lo[r] += sgn hi[nr-1-r]
    for (int e = threadIdx.x; e < nr * NVP; e += GS_OUT * GS_SL) {
      const int r = e / NVP, v = e - r * NVP;
      tile[st][0][r][v] = fma(sgn, tile[st][1][nr - 1 - r][v], tile[st][0][r][v]);
    }
    __syncthreads();      // the fold is visible; everyone is also done with the previous chunk's stage,
    if (threadIdx.x == 0) {                               // which is the one this load refills
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // (it was written by the fold: generic proxy)
      load_chunk(ch + GS_NS - 1);
    }
    // slice sl takes the rows r = sl, sl + GS_SL, ... of the chunk (terms t = ch GS_CH + 1 + r)
#pragma unroll 4
    for (int r = sl; r < nr; r += GS_SL) {
      const double2* lo = reinterpret_cast<const double2*>(&tile[st][0][r][0]);
#pragma unroll
      for (int v2 = 0; v2 < NVP / 2; ++v2) {
        const double2 l2 = lo[v2];
        acc[2 * v2] = fma(s, l2.x, acc[2 * v2]);
        if (2 * v2 + 1 < NV) acc[2 * v2 + 1] = fma(s, l2.y, acc[2 * v2 + 1]);
      }
      const double s2 = fma(s, dc, c * ds), c2 = fma(c, dc, -(s * ds));   // rotate by GS_SL pi a / N
      s = s2; c = c2;
    }
  }
  __syncthreads();        // all slices are done with the ring: `red` aliases it
  if ((N & 1) == 0 && sl == 0 && z == 0) {
    const double sm = sintab[(unsigned)(((unsigned long long)aa * (N / 2)) % N2)];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = fma(sm, x[(size_t)(N / 2 - 1) * NVP + v], acc[v]);
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) red[sl][v][lane] = acc[v];
  __syncthreads();
  // layer l of output `lane` is finished by thread (l, lane)
  double t3[3] = {0.0, 0.0, 0.0};
  if (sl < NL) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      double sum = 0;
#pragma unroll
      for (int q = 0; q < GS_SL; ++q) sum += red[q][3 * sl + i][lane];
      t3[i] = sum;
    }
  }
  if (nz > 1) {
    __shared__ unsigned last_flag;
    const size_t blk = (size_t)bm * gridDim.x + blockIdx.x;
    double* zp = zpart + blk * nz * (NV * GS_OUT);
    if (sl < NL) {
#pragma unroll
      for (int i = 0; i < 3; ++i) zp[(size_t)z * (NV * GS_OUT) + (3 * sl + i) * GS_OUT + lane] = t3[i];
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last_flag = (atomicAdd(&zcount[blk], 1u) == (unsigned)(nz - 1));
    __syncthreads();
    if (!last_flag) return;
    __threadfence();
    if (threadIdx.x == 0) zcount[blk] = 0u;      // ready for the next launch
    if (sl < NL) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        double sum = 0;
        for (int q = 0; q < nz; ++q)
          sum += reinterpret_cast<volatile double*>(zp)[(size_t)q * (NV * GS_OUT) + (3 * sl + i) * GS_OUT + lane];
        t3[i] = sum;
      }
    }
  }
  if (sl < NL && valid)
    border_store<T, STAGE_A>(t3[0], t3[1], t3[2], a, bm, sl, NL, minv, ny, np, n, out, gvec, gvecf, S, bext);
}

// Same transform for short columns (ny <= GS_SMALL_NY, ensembles of small grids): one THREAD per
// (output, layer), so there is no set-up or reduction; the input is a warp-uniform (broadcast)
// load and the rotation advances by pi a / N per term.
template <typename T, bool STAGE_A>
__global__ void __launch_bounds__(128)
border_gsolve_small(const double* __restrict__ in, const double* __restrict__ sintab,
                    const double* __restrict__ minv, int ny, int np, int n, int nl,
                    double* __restrict__ out, double* __restrict__ gvec, float* __restrict__ gvecf,
                    T* __restrict__ S, T* __restrict__ bext, int a_first, int a_count) {
  const int a = a_first + blockIdx.x * blockDim.x + threadIdx.x + 1, plane = blockIdx.y;
  const int bm = plane / nl, l = plane - bm * nl;
  if (a > a_first + a_count) return;
  const unsigned N = ny + 1, N2 = 2 * N;
  const int nvp = border_nvp(nl);
  const double* costab = sintab + N2;
  const double ds = sintab[a], dc = costab[a];        // a < N2
  double s = ds, c = dc, acc0 = 0, acc1 = 0, acc2 = 0;
  const double* x = in + (size_t)bm * ny * nvp + 3 * l;
  const int half = (N - 1) / 2;
  const double sgn = (a & 1) ? 1.0 : -1.0;
  for (int t = 1; t <= half; ++t) {
    const double* lo = x + (size_t)(t - 1) * nvp;
    const double* hi = x + (size_t)(N - t - 1) * nvp;
    acc0 = fma(s, fma(sgn, hi[0], lo[0]), acc0);
    acc1 = fma(s, fma(sgn, hi[1], lo[1]), acc1);
    acc2 = fma(s, fma(sgn, hi[2], lo[2]), acc2);
    const double s2 = fma(s, dc, c * ds), c2 = fma(c, dc, -(s * ds));
    s = s2; c = c2;
  }
  if ((N & 1) == 0) {
    const double sm = sintab[(unsigned)(((unsigned long long)a * (N / 2)) % N2)];
    const double* mid = x + (size_t)(N / 2 - 1) * nvp;
    acc0 = fma(sm, mid[0], acc0);
    acc1 = fma(sm, mid[1], acc1);
    acc2 = fma(sm, mid[2], acc2);
  }
  border_store<T, STAGE_A>(acc0, acc1, acc2, a, bm, l, nl, minv, ny, np, n, out, gvec, gvecf, S, bext);
}

// ------------------------------------------------------------------------------------------
// host side: tables
// ------------------------------------------------------------------------------------------
static int dev_upload(const void* src, size_t bytes, void** dst, size_t* total) {
  SB_CUDA(cudaMalloc(dst, bytes ? bytes : 8));
  if (bytes) SB_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  *total += bytes;
  return 0;
}

static int build_thomas_tables(QgSolver* s, const double* lambdas, int Nx_eig) {
  const int nl = s->nl, ny = s->ny, nc = s->ncols, np = s->np, nstrip = np / SP_W;
  const int hcap = ny - ny / 2;                 // longest run of one half (table indices 0..hcap-1)
  const int m1 = ny / 2;
  std::vector<double> cinf((size_t)nl * np, 0.0), meetc((size_t)nl * 2 * np, 0.0), ctab;
  const int nseg = s->nseg;
  s->seg_len = nseg > 1 ? (((hcap + nseg - 1) / nseg + 63) / 64) * 64 : hcap;
  std::vector<double> segprod(nseg > 1 ? (size_t)4 * nl * nseg * np : 0, 1.0), crow(nseg > 1 ? hcap : 0);
  std::vector<int> Jstrip((size_t)nl * nstrip, 0);
  std::vector<long long> tabOff((size_t)nl * nstrip, 0);
  const double dy2 = s->dy * s->dy;
  s->KB = 0;
  for (int m = 0; m < nl; ++m) {
    s->kbad[m] = 0;
    std::vector<double> delta(np, -4.0);
    std::vector<int> J(np, 0);
    for (int c = 0; c < nc; ++c) {
      const double sn = sin(M_PI * (c + 1) / (2.0 * Nx_eig));
      const double lam_x = -(4.0 / (s->dx * s->dx)) * sn * sn;
      const double eps = (lambdas[m] - lam_x) * dy2;      // |delta| - 2 when the column is definite
      delta[c] = -2.0 - eps;
      int jconv = hcap;
      if (eps > 0.0) {
        // attracting fixed point c* = 2 / (delta - sqrt(delta^2 - 4)), no cancellation; the
        // recurrence approaches it like exp(-2 j theta), theta = acosh(1 + eps/2)
        const double sq = sqrt(eps * (4.0 + eps));
        cinf[(size_t)m * np + c] = 2.0 / (delta[c] - sq);
        const double theta = log1p(0.5 * eps + 0.5 * sq);
        const double jr = log((1.0 - exp(-2.0 * theta)) / 1e-17) / (2.0 * theta) + 2.0;
        if (jr < (double)hcap) jconv = (int)ceil(jr);
      } else {
        s->kbad[m] = std::max(s->kbad[m], c + 1);   // indefinite (oscillatory) column
      }
      J[c] = jconv;
      // coefficients used by the meeting-point solve
      double cj = 0.0, ca = 0.0, cb = 0.0;
      const int ia = m1 - 1, ib = ny - 1 - m1, imax = std::max(ia, ib);
      for (int i = 0; i <= imax && i < jconv; ++i) {
        cj = 1.0 / (delta[c] - cj);
        if (i == ia) ca = cj;
        if (i == ib) cb = cj;
      }
      if (ia >= jconv) ca = cinf[(size_t)m * np + c];
      if (ib >= jconv) cb = cinf[(size_t)m * np + c];
      meetc[((size_t)m * 2 + 0) * np + c] = ca;
      meetc[((size_t)m * 2 + 1) * np + c] = cb;
      if (nseg > 1) {
        // products of -c_i over the segments of both halves, for the elimination (i = s) and the
        // substitution (i = cnt - 1 - s) sequences
        double ci = 0.0;
        for (int i = 0; i < hcap; ++i) { ci = (i < jconv) ? 1.0 / (delta[c] - ci) : cinf[(size_t)m * np + c]; crow[i] = ci; }
        for (int h = 0; h < 2; ++h) {
          const int cnt = h == 0 ? m1 : ny - m1;
          for (int k = 0; k < nseg; ++k) {
            const int sa = std::min(cnt, k * s->seg_len), sb = std::min(cnt, sa + s->seg_len);
            double pe = 1.0, ps = 1.0;
            for (int q = sa; q < sb; ++q) { pe *= -crow[q]; ps *= -crow[cnt - 1 - q]; }
            segprod[((size_t)((0 + h) * nl + m) * nseg + k) * np + c] = pe;
            segprod[((size_t)((2 + h) * nl + m) * nseg + k) * np + c] = ps;
          }
        }
      }
    }
    s->KB = std::max(s->KB, s->kbad[m]);
    for (int st = 0; st < nstrip; ++st) {
      int Js = 0;
      for (int cl = 0; cl < SP_W; ++cl) Js = std::max(Js, J[st * SP_W + cl]);
      Jstrip[(size_t)m * nstrip + st] = Js;
      tabOff[(size_t)m * nstrip + st] = (long long)ctab.size();
      const size_t o = ctab.size();
      ctab.resize(o + (size_t)Js * SP_W, 0.0);
      for (int cl = 0; cl < SP_W; ++cl) {
        const int c = st * SP_W + cl;
        if (c >= nc) continue;
        double cj = 0.0;
        for (int i = 0; i < Js; ++i) {
          // beyond the column's own convergence row keep the exact fixed point
          cj = (i < J[c]) ? 1.0 / (delta[c] - cj) : cinf[(size_t)m * np + c];
          ctab[o + (size_t)i * SP_W + cl] = cj;
        }
      }
    }
  }
  // Launch classes (see ThRows): strips [0, nheavy) form the low-k launch.  A strip is low-k when
  // its coefficient recurrence converges late (many tabulated rows) and, for the fp32 pipeline,
  // when it holds columns k < max(64, n/32) -- near-singular enough to need the fp64 carry -- or
  // an indefinite column.
  s->nheavy = 0;
  for (int m = 0; m < nl; ++m)
    for (int st = 0; st < nstrip; ++st)
      if (Jstrip[(size_t)m * nstrip + st] > 12 * TH_RT) s->nheavy = std::max(s->nheavy, st + 1);
  s->nheavy = std::min(nstrip, std::max(s->nheavy, (s->KB + SP_W - 1) / SP_W));   // indefinite columns
  if (s->dtype == SOMAX_B200_F32) {
    const int klow = std::max(std::max(64, nc / 32), s->KB);
    s->nheavy = std::min(nstrip, std::max(s->nheavy, (klow + SP_W - 1) / SP_W));
    // plain strips read their (few) tabulated rows as floats: repack in place, [i][64] floats
    for (int m = 0; m < nl; ++m)
      for (int st = s->nheavy; st < nstrip; ++st) {
        double* base = ctab.data() + tabOff[(size_t)m * nstrip + st];
        const size_t cnt = (size_t)Jstrip[(size_t)m * nstrip + st] * SP_W;
        float* fb = reinterpret_cast<float*>(base);
        for (size_t i = 0; i < cnt; ++i) { const float v = (float)base[i]; fb[i] = v; }
      }
  }
  if (ctab.empty()) ctab.push_back(0.0);
  if (int rc = dev_upload(ctab.data(), ctab.size() * 8, (void**)&s->ctab, &s->bytes)) return rc;
  if (int rc = dev_upload(Jstrip.data(), Jstrip.size() * 4, (void**)&s->krow, &s->bytes)) return rc;
  {
    // warm-up rows of a thomas_vec_ckpt segment: |c_max|^rows < 1e-13 over the strip's columns
    // (indefinite / near-singular columns never decay: their strips are LOWK and not run by it)
    std::vector<int> vwarm((size_t)nl * nstrip, hcap);
    s->vwarm_max = 0;
    for (int m = 0; m < nl; ++m)
      for (int st = 0; st < nstrip; ++st) {
        double cmax = 0.0;
        for (int cl = 0; cl < SP_W; ++cl) {
          const int c = st * SP_W + cl;
          if (c >= nc) continue;
          const double ci = fabs(cinf[(size_t)m * np + c]);
          cmax = std::max(cmax, ci > 0.0 ? ci : 1.0);
        }
        int w = hcap;
        if (cmax > 0.0 && cmax < 1.0) {
          const double rows = log(1e-13) / log(cmax) + Jstrip[(size_t)m * nstrip + st];
          if (rows < (double)hcap) w = ((int)ceil(rows) + TH_RT - 1) / TH_RT * TH_RT;
        }
        vwarm[(size_t)m * nstrip + st] = w;
        if (st >= s->nheavy) s->vwarm_max = std::max(s->vwarm_max, std::min(w, hcap));   // PLAIN strips only
      }
    if (int rc = dev_upload(vwarm.data(), vwarm.size() * 4, (void**)&s->vwarm, &s->bytes)) return rc;
  }
  if (int rc = dev_upload(tabOff.data(), tabOff.size() * 8, (void**)&s->coff, &s->bytes)) return rc;
  if (int rc = dev_upload(cinf.data(), cinf.size() * 8, (void**)&s->cinf, &s->bytes)) return rc;
  if (int rc = dev_upload(meetc.data(), meetc.size() * 8, (void**)&s->meetc, &s->bytes)) return rc;
  if (nseg > 1) {
    if (int rc = dev_upload(segprod.data(), segprod.size() * 8, (void**)&s->segprod, &s->bytes)) return rc;
    const size_t sbytes = (size_t)s->planes * 2 * nseg * np * 8;
    SB_CUDA(cudaMalloc((void**)&s->segbuf, sbytes));
    SB_CUDA(cudaMemset(s->segbuf, 0, sbytes));
    s->bytes += sbytes;
  }
  if (s->dtype == SOMAX_B200_F32 && s->kind == SOMAX_B200_SOLVER_FFT) {
    s->nck = (hcap + TH_RT - 1) / TH_RT;
    const size_t cb = (size_t)s->planes * nstrip * 2 * s->nck * SP_W * sizeof(float);
    SB_CUDA(cudaMalloc((void**)&s->ckpt, cb));
    SB_CUDA(cudaMemset(s->ckpt, 0, cb));
    s->bytes += cb;
  }
  {
    size_t mb = (size_t)s->planes * 2 * np * 8;
    SB_CUDA(cudaMalloc((void**)&s->meet, mb));
    SB_CUDA(cudaMemset(s->meet, 0, mb));
    SB_CUDA(cudaMalloc((void**)&s->meet1, mb));
    SB_CUDA(cudaMemset(s->meet1, 0, mb));
    s->bytes += 2 * mb;
    const size_t pb = (size_t)s->planes * 4 * nstrip * ny * (s->dtype == SOMAX_B200_F32 ? 4 : 8);
    SB_CUDA(cudaMalloc(&s->part, pb));
    SB_CUDA(cudaMemset(s->part, 0, pb));
    s->bytes += pb;
  }
  {
    s->KBs = std::max(2, (s->KB + 1) & ~1);      // side-buffer row length: even (16-byte bulk copies)
    size_t nb = (size_t)s->planes * ny * s->KBs * 8;
    SB_CUDA(cudaMalloc((void**)&s->dbad, nb));
    SB_CUDA(cudaMemset(s->dbad, 0, nb));
    SB_CUDA(cudaMalloc((void**)&s->dbad1, nb));
    SB_CUDA(cudaMemset(s->dbad1, 0, nb));
    s->bytes += 2 * nb;
  }
  return 0;
}

template <typename T>
static int build_fft_tables(QgSolver* s, const double* lambdas) {
  const int n = s->nx, ny = s->ny, nl = s->nl, nc = s->ncols;
  const double b = 1.0 / (s->dx * s->dx);
  std::vector<C2<T>> tw(2 * (size_t)n);
  for (int t = 0; t < 2 * n; ++t) {
    double a = -M_PI * t / n;
    tw[t].x = (T)cos(a); tw[t].y = (T)sin(a);
  }
  if (int rc = dev_upload(tw.data(), tw.size() * sizeof(C2<T>), &s->tw, &s->bytes)) return rc;
  {
    // compact per-pass butterfly twiddles (layout: fft.cuh twc_offset)
    std::vector<C2<T>> twc;
    int lgLc = s->plan.lgn;
    for (int ps = 0; ps < s->plan.npass; ++ps) {
      const int lr = s->plan.lgr[ps], Lc = 1 << lgLc, M = Lc >> lr;
      for (int pos = 0; pos < M; ++pos)
        for (int q = 1; q <= 4; q *= 2) {
          const double a = -2.0 * M_PI * (double)q * (double)pos / (double)Lc;
          twc.push_back({(T)cos(a), (T)sin(a)});
        }
      lgLc -= lr;
    }
    if (int rc = dev_upload(twc.data(), twc.size() * sizeof(C2<T>), &s->twc, &s->bytes)) return rc;
    if (sizeof(T) == 4 && s->plan.lgn >= 12 && s->plan.lgn <= 14 && !getenv("SOMAX_B200_FFT_RADIX8")) {
      // three-pass kernel (BigCfg): powers W^(2^jj) of the pass-1 and pass-2 twiddles
      const int lgn = s->plan.lgn, lg1 = lgn >= 13 ? 5 : 4, lg2 = lgn >= 14 ? 5 : 4, r3 = 16;
      const int G = n >> lg1;
      std::vector<C2<T>> twb;
      for (int jj = 0; jj < lg1; ++jj)
        for (int lt = 0; lt < G; ++lt) {
          const double a = -2.0 * M_PI * (double)(((long long)lt << jj) % n) / (double)n;
          twb.push_back({(T)cos(a), (T)sin(a)});
        }
      for (int jj = 0; jj < lg2; ++jj)
        for (int pos = 0; pos < r3; ++pos) {
          const double a = -2.0 * M_PI * (double)((pos << jj) % G) / (double)G;
          twb.push_back({(T)cos(a), (T)sin(a)});
        }
      if (int rc = dev_upload(twb.data(), twb.size() * sizeof(C2<T>), &s->twb, &s->bytes)) return rc;
    }
  }
  // block column 1 weights sin(pi k / n); block column n-1 is the same times (-1)^(k+1)
  std::vector<double> sig(nc), lamx(nc), bsig(nc), sig2n(nc);
  for (int c = 0; c < nc; ++c) {
    const int k = c + 1;
    sig[c] = sin(M_PI * k / n);
    const double sn = sin(M_PI * k / (2.0 * n));
    lamx[c] = -(4.0 * b) * sn * sn;
    bsig[c] = b * sig[c];
    sig2n[c] = (2.0 / n) * sig[c];
  }
  // Schur block of the border unknowns (g0, g1, g2) per (mode, y-mode), inverted here:
  //   [ d - b^2 al   -b^2 be      0 ]        al = (2/n) sum_k sig_k^2 / (lamx_k + mu)
  //   [ -b^2 be      d - b^2 al   b ]        be = (2/n) sum_k (-1)^(k+1) sig_k^2 / (lamx_k + mu)
  //   [ 0            b            d ]        d  = mu - 2 b,  mu = lamy_l - lambda_m
  std::vector<double> minv((size_t)nl * ny * 9);
  for (int m = 0; m < nl; ++m)
    for (int l = 1; l <= ny; ++l) {
      const double sn = sin(M_PI * l / (2.0 * (ny + 1)));
      const double mu = -(4.0 / (s->dy * s->dy)) * sn * sn - lambdas[m];
      double al = 0, be = 0;
      for (int c = 0; c < nc; ++c) {
        const double t = sig[c] * sig[c] / (lamx[c] + mu);
        al += t; be += (c & 1) ? -t : t;
      }
      al *= 2.0 / n; be *= 2.0 / n;
      const double d = mu - 2.0 * b, A = d - b * b * al, B = -b * b * be;
      const double M[3][3] = {{A, B, 0.0}, {B, A, b}, {0.0, b, d}};
      const double det = M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) -
                         M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
                         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
      double* o = minv.data() + ((size_t)m * ny + (l - 1)) * 9;
      for (int i = 0; i < 3; ++i)
        for (int jj = 0; jj < 3; ++jj) {
          const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (jj + 1) % 3, j2 = (jj + 2) % 3;
          // inverse = adjugate / det; cofactor (jj, i) by cyclic indices
          o[i * 3 + jj] = (M[j1][i1] * M[j2][i2] - M[j1][i2] * M[j2][i1]) / det;
        }
    }
  // sin(pi t / N), t < 2N, followed by cos(pi t / N)
  const size_t N2 = 2 * (size_t)(ny + 1);
  std::vector<double> sintab(2 * N2);
  for (size_t t = 0; t < N2; ++t) {
    sintab[t] = sin(M_PI * (double)t / (ny + 1));
    sintab[N2 + t] = cos(M_PI * (double)t / (ny + 1));
  }
  if (int rc = dev_upload(bsig.data(), nc * 8, (void**)&s->bsig, &s->bytes)) return rc;
  {
    std::vector<T> sg((size_t)s->np, (T)0);      // working precision, zero-padded to np
    for (int c = 0; c < nc; ++c) sg[c] = (T)sig2n[c];
    if (int rc = dev_upload(sg.data(), sg.size() * sizeof(T), &s->sig2n, &s->bytes)) return rc;
  }
  if (int rc = dev_upload(minv.data(), minv.size() * 8, (void**)&s->minv, &s->bytes)) return rc;
  if (int rc = dev_upload(sintab.data(), sintab.size() * 8, (void**)&s->sintab, &s->bytes)) return rc;
  size_t vb = (size_t)s->planes * ny * 8;
  const size_t rb = (size_t)s->batch * ny * border_nvp(nl) * 8;      // <= 4 vb
  SB_CUDA(cudaMalloc((void**)&s->rvec, rb));
  SB_CUDA(cudaMemset(s->rvec, 0, rb));
  SB_CUDA(cudaMalloc((void**)&s->ghat, rb));
  SB_CUDA(cudaMemset(s->ghat, 0, rb));
  SB_CUDA(cudaMalloc((void**)&s->gvec, 2 * vb));
  SB_CUDA(cudaMalloc((void**)&s->gvecf, vb));
  if (!(ny <= GS_SMALL_NY && (long)s->planes * ny >= 65536)) {      // the blocked border transform is in use
    const size_t nblk = (size_t)s->batch * (2 * ((ny + 2 * GS_OUT - 1) / (2 * GS_OUT)));
    const size_t zb = nblk * GS_ZMAX * 12 * GS_OUT * sizeof(double);
    SB_CUDA(cudaMalloc((void**)&s->zpart, zb));
    SB_CUDA(cudaMalloc((void**)&s->zcount, nblk * sizeof(unsigned)));
    SB_CUDA(cudaMemset(s->zcount, 0, nblk * sizeof(unsigned)));
    s->bytes += zb + nblk * sizeof(unsigned);
  }
  SB_CUDA(cudaMalloc(&s->bext, 3 * (size_t)s->planes * ny * sizeof(T)));
  SB_CUDA(cudaMemset(s->bext, 0, 3 * (size_t)s->planes * ny * sizeof(T)));
  s->bytes += 2 * rb + 3 * vb + 2 * (size_t)s->planes * ny * sizeof(T);
  return 0;
}

template <typename T>
static int build_dense_tables(QgSolver* s) {
  const int n = s->nx + 2;      // every column of the array is an unknown
  std::vector<T> mat((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k)
      mat[(size_t)i * n + k] = (T)sin(M_PI * (double)(i + 1) * (double)(k + 1) / (n + 1));
  return dev_upload(mat.data(), mat.size() * sizeof(T), &s->dstmat, &s->bytes);
}

int qg_solver_create(QgSolver** out, int dtype, int batch, int nl, int ny, int nx, double dx,
                     double dy, const double* Cl2m, const double* Cm2l, const double* lambdas,
                     int solver_kind, int nseg, int rows, int jo, int ylo, int yhi) {
  *out = nullptr;
  if (nl < 1 || nl > QG_MAX_NL) return fail(SOMAX_B200_ERR_UNSUPPORTED, "QG supports 1 <= nl <= 4");
  const bool pow2 = nx >= 8 && (nx & (nx - 1)) == 0;
  int kind = solver_kind;
  if (kind == SOMAX_B200_SOLVER_AUTO) kind = pow2 ? SOMAX_B200_SOLVER_FFT : SOMAX_B200_SOLVER_DENSE;
  const size_t es = dtype == SOMAX_B200_F32 ? 4 : 8;
  if (kind == SOMAX_B200_SOLVER_FFT) {
    if (!pow2) return fail(SOMAX_B200_ERR_UNSUPPORTED, "FFT solver needs nx = 2^p >= 8");
    if ((size_t)fft_padded_len(nx) * 2 * es > 200 * 1024)
      return fail(SOMAX_B200_ERR_UNSUPPORTED, "FFT solver: one row must fit in shared memory (nx <= 8192 fp64 / 16384 fp32)");
  } else if (kind == SOMAX_B200_SOLVER_DENSE) {
    if (nx > 2046) return fail(SOMAX_B200_ERR_UNSUPPORTED, "dense DST solver limited to nx <= 2046; use nx = 2^p for the FFT path");
  } else {
    return fail(SOMAX_B200_ERR_INVALID, "unknown solver kind");
  }
  auto* s = new QgSolver();
  s->dtype = dtype; s->batch = batch; s->nl = nl; s->nx = nx; s->kind = kind;
  s->dx = dx; s->dy = dy; s->L = make_layout(batch, nl, ny, nx);
  s->ny = rows > 0 ? rows : ny + 2; s->jo = rows > 0 ? jo : 0;
  s->ylo = rows > 0 ? ylo : 1; s->yhi = rows > 0 ? yhi : 1;
  ny = s->ny;      // from here on: solver rows
  // nseg = 0 (a whole grid on one device): the PLAIN sweeps fill the GPU unsegmented, but the
  // handful of LOWK CTAs are serial chains over ny/2 rows that end up as the critical path of both
  // solves (0.13 + 0.40 ms against 0.08 + 0.29 ms of the PLAIN class at 8192^2), so THEY are cut
  // into 8 segments once the columns are long
  s->seg_plain = nseg != 0;
  if (nseg == 0) nseg = ny + 2 >= 2048 ? 8 : 1;
  if (const char* e = getenv("SOMAX_B200_LOWK_NSEG")) { if (!s->seg_plain) nseg = atoi(e); }
  s->nseg = std::max(1, std::min(nseg, 16));
  s->ncols = (kind == SOMAX_B200_SOLVER_FFT) ? nx - 1 : nx + 2;
  s->np = ((std::max(nx, s->ncols) + SP_W - 1) / SP_W) * SP_W; s->planes = batch * nl;
  for (int a = 0; a < QG_MAX_NL; ++a)
    for (int c = 0; c < QG_MAX_NL; ++c) {
      s->l2m.c[a][c] = (a < nl && c < nl) ? Cl2m[a * nl + c] : 0.0;
      s->m2l.c[a][c] = (a < nl && c < nl) ? Cm2l[a * nl + c] : 0.0;
    }
  // the reference mixes in working precision (einsum on fp32 arrays, core/transforms.py:218-224)
  if (dtype == SOMAX_B200_F32)
    for (int a = 0; a < QG_MAX_NL; ++a)
      for (int c = 0; c < QG_MAX_NL; ++c) {
        s->l2m.c[a][c] = (double)(float)s->l2m.c[a][c];
        s->m2l.c[a][c] = (double)(float)s->m2l.c[a][c];
      }
  s->plan = make_fft_plan(nx);
  int rc = 0;
  if (cudaStreamCreateWithFlags(&s->aux, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    delete s;
    return fail(SOMAX_B200_ERR_CUDA, "stream/event creation failed");
  }
  const size_t sb_ = (size_t)s->planes * ny * s->np * es;
  auto alloc0 = [&](void** p) -> int {
    cudaError_t e = cudaMalloc(p, sb_);
    if (e != cudaSuccess) return fail(SOMAX_B200_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cudaMemset(*p, 0, sb_);
    s->bytes += sb_;
    return 0;
  };
  rc = alloc0(&s->S);
  if (!rc && kind == SOMAX_B200_SOLVER_FFT) rc = alloc0(&s->W);
  if (!rc) rc = build_thomas_tables(s, lambdas, kind == SOMAX_B200_SOLVER_FFT ? nx : nx + 3);
  if (!rc) {
    if (kind == SOMAX_B200_SOLVER_FFT)
      rc = dtype == SOMAX_B200_F32 ? build_fft_tables<float>(s, lambdas) : build_fft_tables<double>(s, lambdas);
    else
      rc = dtype == SOMAX_B200_F32 ? build_dense_tables<float>(s) : build_dense_tables<double>(s);
  }
  if (rc) { qg_solver_destroy(s); return rc; }
  *out = s;
  return 0;
}

void qg_solver_destroy(QgSolver* s) {
  if (!s) return;
  void* ptrs[] = {s->zpart, s->zcount, s->vwarm, s->pvec, s->ckpt, s->segbuf, s->segprod, s->S, s->W, s->ctab, s->coff, s->krow, s->cinf, s->dbad, s->dbad1, s->meet1, s->part, s->bsig, s->sig2n,
                  s->minv, s->bext, s->sintab, s->rvec, s->ghat, s->gvec, s->gvecf, s->tw, s->twc, s->twb, s->dstmat, s->meet, s->meetc};
  for (void* p : ptrs) cudaFree(p);
  if (s->aux) cudaStreamDestroy(s->aux);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_join) cudaEventDestroy(s->ev_join);
  delete s;
}

size_t qg_solver_bytes(const QgSolver* s) { return s ? s->bytes : 0; }

void qg_solver_set_scatter(QgSolver* s, void* const* peerS, int nranks, int spr, int row0, int ny_cols) {
  for (int r = 0; r < 16; ++r) s->sc_peer[r] = r < nranks ? peerS[r] : nullptr;
  int lg = 0;
  while ((1 << lg) < spr) ++lg;
  s->sc_lgspr = lg; s->sc_row0 = row0; s->sc_ny = ny_cols;
}

int qg_solver_set_rank_reduce(QgSolver* s, int nranks, int me, int s0, int s1) {
  const size_t nb = (size_t)nranks * s->planes * 2 * s->ny * sizeof(double);
  SB_CUDA(cudaMalloc((void**)&s->pvec, nb));
  SB_CUDA(cudaMemset(s->pvec, 0, nb));
  s->bytes += nb;
  s->pv_n = nranks; s->pv_me = me; s->pv_s0 = s0; s->pv_s1 = s1;
  return 0;
}

void qg_solver_set_push(QgSolver* s, void* const* peerR, int nranks, const int* row0) {
  for (int r = 0; r < 16; ++r) s->ps_peer[r] = r < nranks ? peerR[r] : nullptr;
  for (int r = 0; r <= nranks; ++r) s->ps_row0[r] = row0[r];
  s->ps_n = nranks;
}
int qg_solver_kind(const QgSolver* s) { return s->kind; }

template <typename T, bool SUBST, bool FROM_VEC, int KIND, bool TAB>
static int launch_thomas_one(const char* tag, const ThomasTab& tb, int strip_first, int nstrips, int planes,
                             const T* in, const T* V, const double* gvec, const float* gvecf,
                             const double* bsig, T* out, cudaStream_t st) {
  if (nstrips <= 0) return 0;
  if (!TAB && tb.nseg > 1 && !tb.seg_plain) {       // this class runs unsegmented
    ThomasTab t0 = tb;
    t0.nseg = 1; t0.seg_plain = true;
    return launch_thomas_one<T, SUBST, FROM_VEC, KIND, TAB>(tag, t0, strip_first, nstrips, planes, in, V, gvec, gvecf, bsig, out, st);
  }
  constexpr int NS = ThStages<T, KIND, TAB>::v;
  constexpr int RT = ThRows<T, KIND, TAB>::v;
  constexpr int NOP = (KIND == 2 && (TAB || sizeof(T) == 8)) ? 2 : 1;      // operand tiles (and side tiles) per stage
  constexpr size_t smem0 = (size_t)NS * RT * TH_COLS * ((TAB ? sizeof(double) : 0) + sizeof(T) * NOP) + NS * 8;
  static_assert(smem0 <= 227 * 1024, "sweep ring exceeds shared memory");
  const size_t smem = smem0 + (TAB ? (size_t)NS * RT * tb.KB * sizeof(double) * NOP : 0);
  if (smem > 227 * 1024)
    return fail(SOMAX_B200_ERR_UNSUPPORTED, "too many indefinite Helmholtz columns for the staged side buffer");
  if (int rc = ensure_dyn_smem((const void*)thomas_sweep<T, SUBST, FROM_VEC, KIND, TAB>, smem)) return rc;
  if (tb.nseg > 1) {
    ThomasTab t1 = tb;
    t1.pass = 1;
    prof_begin(TAB ? "thomas_probe_lowk" : "thomas_probe", st);
    thomas_sweep<T, SUBST, FROM_VEC, KIND, TAB><<<dim3(nstrips, planes, 2 * tb.nseg), TH_THREADS, smem, st>>>(
        t1, strip_first, in, V, gvec, gvecf, bsig, out);
    SB_LAUNCH_CHECK();
    t1.pass = 2;
    prof_begin(tag, st);
    thomas_sweep<T, SUBST, FROM_VEC, KIND, TAB><<<dim3(nstrips, planes, 2 * tb.nseg), TH_THREADS, smem, st>>>(
        t1, strip_first, in, V, gvec, gvecf, bsig, out);
    SB_LAUNCH_CHECK();
    return 0;
  }
  prof_begin(tag, st);
  thomas_sweep<T, SUBST, FROM_VEC, KIND, TAB><<<dim3(nstrips, planes, 2), TH_THREADS, smem, st>>>(
      tb, strip_first, in, V, gvec, gvecf, bsig, out);
  SB_LAUNCH_CHECK();
  return 0;
}

// One solve = elimination + substitution sweeps.  Strips are independent of each other, so the
// two launch classes run as two chains on two streams and join after the pair.
//   PHASE 0: S <- A^-1 S (plain solve; dense-x path).
//   PHASE 1: S <- eliminated S (D), border row sums of A^-1 S into tb.part (x itself is not stored).
//   PHASE 2: W <- eliminated (gvec x 1) (E), then S <- back-substitution of D - bsig * E.
template <typename T, int PHASE>
static int launch_solve(QgSolver* s, const ThomasTab& tb, T* S, T* W, cudaStream_t st, int sa = 0, int sb = -1) {
  // strips [sa, sb) only (slab-distributed solve: every rank owns a range of x-wavenumber strips)
  const int nstrip_all = s->np / SP_W;
  if (sb < 0) sb = nstrip_all;
  const int hA = std::min(sa, s->nheavy), hB = std::min(sb, s->nheavy);   // LOWK strips [hA, hB)
  const int pA = std::max(sa, s->nheavy), pB = std::max(sb, s->nheavy);   // PLAIN strips [pA, pB)
  const int nh = hB - hA, npl = pB - pA;
  constexpr bool SECOND = PHASE == 2;
  constexpr int KIND = PHASE;
  const char* tf = SECOND ? "thomas_fwd_1" : "thomas_fwd_0";
  const char* tbk = SECOND ? "thomas_bwd_1" : "thomas_bwd_0";
  const char* tfl = SECOND ? "thomas_fwd_1_lowk" : "thomas_fwd_0_lowk";
  const char* tbl = SECOND ? "thomas_bwd_1_lowk" : "thomas_bwd_0_lowk";
  const T* fin = SECOND ? nullptr : S;      // elimination input
  T* fout = SECOND ? W : S;                 // eliminated right-hand side
  const T* bin = S;                         // substitution input (D)
  const T* Vv = SECOND ? W : nullptr;       // second operand (E)
  const double* bs = SECOND ? s->bsig : nullptr;
  // The LOWK launches go first, on the caller's stream, so that their few long CTAs are resident
  // before the wide PLAIN launches (auxiliary stream, released by an event a few microseconds
  // later) fill every SM's shared memory; measured the other way round, the LOWK CTAs could not be
  // placed until the PLAIN kernel drained and the two chains ran back to back.
  static const bool serial = getenv("SOMAX_B200_SERIAL_SWEEPS") != nullptr;      // (diagnostic: both classes on one stream)
  const bool two = nh > 0 && npl > 0 && !serial;
  cudaStream_t pl = two ? s->aux : st;
  if (two) {
    SB_CUDA(cudaEventRecord(s->ev_fork, st));
    SB_CUDA(cudaStreamWaitEvent(s->aux, s->ev_fork, 0));
  }
  if (int rc = launch_thomas_one<T, false, SECOND, 0, true>(tfl, tb, hA, nh, s->planes, fin, nullptr, s->gvec, s->gvecf, nullptr, fout, st)) return rc;
  if constexpr (SECOND && sizeof(T) == 4) {
    // PLAIN fp32 strips: checkpoints of the eliminated border right-hand side, nothing stored
    if (npl > 0) {
      const int hcap = s->ny - s->ny / 2;
      const int vseg = hcap >= 1024 ? VC_NSEG : 1;
      // a segment plus its warm-up (6 CTAs / SM and six waves when sized for a whole half)
      const int seg_rows = TH_RT * ((hcap / TH_RT + vseg - 1) / vseg + 1) + TH_RT;
      const size_t smem = (size_t)2 * std::min(hcap, seg_rows + s->vwarm_max) * sizeof(float);
      if (int rc = ensure_dyn_smem((const void*)thomas_vec_ckpt, smem)) return rc;
      prof_begin(tf, pl);
      thomas_vec_ckpt<<<dim3(npl, s->planes, 2 * vseg), TH_COLS, smem, pl>>>(tb, pA, s->gvecf, vseg);
      SB_LAUNCH_CHECK();
    }
  } else {
    if (int rc = launch_thomas_one<T, false, SECOND, 0, false>(tf, tb, pA, npl, s->planes, fin, nullptr, s->gvec, s->gvecf, nullptr, fout, pl)) return rc;
  }
  if (int rc = launch_thomas_one<T, true, false, KIND, true>(tbl, tb, hA, nh, s->planes, bin, Vv, nullptr, nullptr, bs, S, st)) return rc;
  // (the PLAIN fp32 KIND 2 sweep recomputes E from the shared right-hand side: it reads gvecf)
  if (int rc = launch_thomas_one<T, true, false, KIND, false>(tbk, tb, pA, npl, s->planes, bin, Vv, s->gvec, s->gvecf, bs, S, pl)) return rc;
  if (two) {
    SB_CUDA(cudaEventRecord(s->ev_join, s->aux));
    SB_CUDA(cudaStreamWaitEvent(st, s->ev_join, 0));
  }
  return 0;
}

static ThomasTab make_tab(const QgSolver* s) {
  ThomasTab tb;
  tb.ctabB = s->ctab; tb.tabOff = s->coff; tb.Jstrip = s->krow; tb.cinf = s->cinf; tb.meetc = s->meetc;
  tb.nstrip = s->np / SP_W;
  for (int m = 0; m < QG_MAX_NL; ++m) tb.kbad[m] = s->kbad[m];
  tb.KB = s->KBs; tb.dbad = s->dbad; tb.dbad1 = s->dbad1; tb.meet = s->meet; tb.meet1 = s->meet1;
  tb.part = s->part; tb.sig2n = s->sig2n; tb.ny = s->ny; tb.np = s->np; tb.ncols = s->ncols; tb.nl = s->nl;
  tb.dy2 = s->dy * s->dy;
  tb.nseg = s->nseg; tb.seg_len = s->seg_len; tb.pass = 0; tb.segbuf = s->segbuf; tb.segprod = s->segprod;
  tb.seg_plain = s->seg_plain;
  tb.ckpt = s->ckpt; tb.nck = s->nck; tb.vwarm = s->vwarm;
  for (int r = 0; r < 16; ++r) tb.ppeer[r] = s->ps_peer[r];
  for (int r = 0; r < 17; ++r) tb.prow[r] = s->ps_row0[r];
  tb.pn = s->ps_n;
  return tb;
}

template <typename T>
static void make_row_args(const QgSolver* s, RowArgsCT<T>& Af, RowArgsCT<T>& Ai) {
  Af.L = s->L; Af.ny = s->ny; Af.np = s->np; Af.nl = s->nl; Af.nrows = s->batch * s->ny;
  Af.jo = s->jo; Af.ylo = s->ylo; Af.yhi = s->yhi; Af.ringmode = 0; Af.bext = (T*)s->bext;
  for (int r = 0; r < 16; ++r) Af.peer[r] = (T*)s->sc_peer[r];
  Af.lgspr = s->sc_lgspr; Af.prow0 = s->sc_row0; Af.pny = s->sc_ny;
  Af.tw = (const C2<T>*)s->tw; Af.twc = (const C2<T>*)s->twc; Af.twb = (const C2<T>*)s->twb; Af.scale = (T)1;
  Ai = Af; Ai.scale = (T)(2.0 / s->nx);
  for (int a = 0; a < QG_MAX_NL; ++a)
    for (int c = 0; c < QG_MAX_NL; ++c) { Af.mix[a][c] = (T)s->l2m.c[a][c]; Ai.mix[a][c] = (T)s->m2l.c[a][c]; }
}

// The FFT-path inversion in its four stages.  qg_solver_run chains them on one device; the
// slab-distributed model (qg_slab.cuh) runs the row stages on a y-slab and the column stages on
// a range of x-wavenumber strips, with peer-memory exchanges in between.
template <typename T>
int qg_solver_rows_fwd(QgSolver* s, const T* q, int ring_zero, cudaStream_t st) {
  RowArgsCT<T> Af, Ai;
  make_row_args<T>(s, Af, Ai);
  Af.ringmode = ring_zero;
  return launch_rowdst<T, false>(s->plan.lgn, Af, q, (T*)s->S, st);
}

template <typename T>
int qg_solver_rows_inv(QgSolver* s, T* psi, int keep_ring, cudaStream_t st) {
  RowArgsCT<T> Af, Ai;
  make_row_args<T>(s, Af, Ai);
  Ai.ringmode = !keep_ring;
  return launch_rowdst<T, true>(s->plan.lgn, Ai, (const T*)s->S, psi, st);
}

template <typename T>
int qg_solver_cols(QgSolver* s, int phase, int sa, int sb, cudaStream_t st) {
  const ThomasTab tb = make_tab(s);
  if (phase == 1) return launch_solve<T, 1>(s, tb, (T*)s->S, nullptr, st, sa, sb);
  return launch_solve<T, 2>(s, tb, (T*)s->S, (T*)s->W, st, sa, sb);
}

// stage 0: reduction of the border partials (rvec); stage 1: DST + Schur diagonal (ghat) for the
// outputs [a0, a1); stage 2: second DST (gvec, gvecf, border column of S) for the outputs [a0, a1).
// The slab-distributed model gives every rank a slice of the outputs and exchanges the slices.
template <typename T>
int qg_solver_border_stage(QgSolver* s, int stage, int a0, int a1, cudaStream_t st) {
  const int ny = s->ny, n = s->nx, np = s->np, nl = s->nl;
  T* S = (T*)s->S;
  if (a1 < 0) a1 = ny;
  const int cnt = a1 - a0;
  if (stage == -1) {      // slab model: pre-sum the partials of this rank's strips into its slot of pvec
    if (!s->pvec) return fail(SOMAX_B200_ERR_INVALID, "rank reduction not configured");
    prof_begin("border_presum", st);
    border_presum<T><<<dim3((ny + 31) / 32, s->planes * 2), 256, 0, st>>>(
        (const T*)s->part, ny, 2 * (np / SP_W), 2 * s->pv_s0, 2 * s->pv_s1,
        s->pvec + (size_t)s->pv_me * s->planes * 2 * ny);
    SB_LAUNCH_CHECK();
    return 0;
  }
  if (stage == 0 && s->pvec) {
    const double b = 1.0 / (s->dx * s->dx);
    prof_begin("border_reduce", st);
    border_reduce_ranks<T><<<dim3((ny + 255) / 256, s->planes), 256, 0, st>>>(
        (const T*)s->bext, s->pvec, s->pv_n, s->planes, ny, nl, b, s->rvec);
    SB_LAUNCH_CHECK();
    return 0;
  }
  if (stage == 0) {
    const double b = 1.0 / (s->dx * s->dx);
    prof_begin("border_reduce", st);
    border_reduce<T><<<dim3((ny + 31) / 32, s->planes), 256, 0, st>>>(S, (const T*)s->bext, (const T*)s->part, ny, np, s->ncols, 2 * (np / SP_W), nl, b, s->rvec);
    SB_LAUNCH_CHECK();
    return 0;
  }
  if (cnt <= 0) return 0;
  // one thread per (output, layer) pays off only when there are enough of them to fill the GPU
  // (ensembles of small grids); a single grid runs the blocked transform
  const bool small = ny <= GS_SMALL_NY && (long)s->planes * ny >= 65536;
  const double* in = stage == 1 ? s->rvec : s->ghat;
  prof_begin(stage == 1 ? "border_gsolve_a" : "border_gsolve_b", st);
  if (small) {
    const dim3 g((cnt + 127) / 128, s->planes);
    if (stage == 1) border_gsolve_small<T, true><<<g, 128, 0, st>>>(in, s->sintab, s->minv, ny, np, n, nl, s->ghat, nullptr, nullptr, nullptr, nullptr, a0, cnt);
    else border_gsolve_small<T, false><<<g, 128, 0, st>>>(in, s->sintab, s->minv, ny, np, n, nl, nullptr, s->gvec, s->gvecf, S, (T*)s->bext, a0, cnt);
  } else {
    const int nblk = 2 * ((cnt + 2 * GS_OUT - 1) / (2 * GS_OUT));      // blocks of 32 outputs of one parity
    const int nchunk_all = ((ny + 1 - 1) / 2 + GS_CH - 1) / GS_CH;
    // enough CTAs for two per SM: split the terms when there are few output blocks
    int nz = (2 * 148 + nblk * s->batch - 1) / (nblk * s->batch);
    nz = std::max(1, std::min(std::min(nz, GS_ZMAX), nchunk_all));
    const dim3 g(nblk, s->batch, nz);
#define SB_BDST(NLV)                                                                                         \
    if (stage == 1) border_dst<T, NLV, true><<<g, GS_OUT * GS_SL, 0, st>>>(in, s->sintab, s->minv, ny, np, n, s->ghat, nullptr, nullptr, nullptr, nullptr, a0, cnt, s->zpart, s->zcount); \
    else border_dst<T, NLV, false><<<g, GS_OUT * GS_SL, 0, st>>>(in, s->sintab, s->minv, ny, np, n, nullptr, s->gvec, s->gvecf, S, (T*)s->bext, a0, cnt, s->zpart, s->zcount)
    switch (nl) {
      case 1: SB_BDST(1); break;
      case 2: SB_BDST(2); break;
      case 3: SB_BDST(3); break;
      default: SB_BDST(4); break;
    }
#undef SB_BDST
  }
  SB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int qg_solver_border(QgSolver* s, cudaStream_t st) {
  for (int stage = 0; stage < 3; ++stage)
    if (int rc = qg_solver_border_stage<T>(s, stage, 0, -1, st)) return rc;
  return 0;
}

template <typename T>
int qg_solver_run(QgSolver* s, const T* q, T* psi, int ring_zero, int keep_ring, cudaStream_t st) {
  const int ny = s->ny, np = s->np, nl = s->nl;
  if (s->kind == SOMAX_B200_SOLVER_FFT) {
    if (int rc = qg_solver_rows_fwd<T>(s, q, ring_zero, st)) return rc;
    if (int rc = qg_solver_cols<T>(s, 1, 0, -1, st)) return rc;
    if (int rc = qg_solver_border<T>(s, st)) return rc;
    if (int rc = qg_solver_cols<T>(s, 2, 0, -1, st)) return rc;
    if (int rc = qg_solver_rows_inv<T>(s, psi, keep_ring, st)) return rc;
  } else {
    const ThomasTab tb = make_tab(s);
    T* S = (T*)s->S;
    const int n = s->nx + 2;
    const size_t smem = (size_t)nl * n * sizeof(T);
    if (int rc = ensure_dyn_smem((const void*)rowdst_dense<T, false>, smem)) return rc;
    if (int rc = ensure_dyn_smem((const void*)rowdst_dense<T, true>, smem)) return rc;
    const int threads = std::min(256, ((n + 31) / 32) * 32);
    prof_begin("rowdst_dense_0", st);
    rowdst_dense<T, false><<<s->batch * ny, threads, smem, st>>>(s->L, ny, n, np, nl, ring_zero, s->l2m, (const T*)s->dstmat, q, S, 1.0);
    SB_LAUNCH_CHECK();
    if (int rc = launch_solve<T, 0>(s, tb, S, nullptr, st)) return rc;
    prof_begin("rowdst_dense_1", st);
    rowdst_dense<T, true><<<s->batch * ny, threads, smem, st>>>(s->L, ny, n, np, nl, !keep_ring, s->m2l, (const T*)s->dstmat, S, psi, 2.0 / (n + 1));
    SB_LAUNCH_CHECK();
  }
  return 0;
}

QgSolverView qg_solver_view(const QgSolver* s) {
  QgSolverView v;
  v.S = s->S; v.part = s->part; v.pvec = s->pvec; v.bext = s->bext; v.ghat = s->ghat; v.gvec = s->gvec; v.gvecf = s->gvecf; v.ny = s->ny; v.nx = s->nx; v.np = s->np; v.planes = s->planes;
  v.nstrip = s->np / SP_W; v.ncols = s->ncols; v.kind = s->kind; v.nheavy = s->nheavy;
  return v;
}

#if defined(SB_TH_DEBUG) || defined(SB_TH_PHASES)
extern "C" int somax_b200_debug_dump(unsigned long long* out, unsigned* n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(n, g_dbg_n, sizeof(unsigned));
  cudaMemcpyFromSymbol(out, g_dbg, sizeof(unsigned long long) * 4096 * 4);
  unsigned zero = 0;
  cudaMemcpyToSymbol(g_dbg_n, &zero, sizeof(unsigned));
  return 0;
}
#endif

template int qg_solver_run<float>(QgSolver*, const float*, float*, int, int, cudaStream_t);
template int qg_solver_run<double>(QgSolver*, const double*, double*, int, int, cudaStream_t);
#define SB_INST_STAGES(T)                                                  \
  template int qg_solver_rows_fwd<T>(QgSolver*, const T*, int, cudaStream_t);   \
  template int qg_solver_rows_inv<T>(QgSolver*, T*, int, cudaStream_t);         \
  template int qg_solver_cols<T>(QgSolver*, int, int, int, cudaStream_t);  \
  template int qg_solver_border<T>(QgSolver*, cudaStream_t);                \
  template int qg_solver_border_stage<T>(QgSolver*, int, int, int, cudaStream_t);
SB_INST_STAGES(float)
SB_INST_STAGES(double)

}  // namespace sb

// ------------------------------------------------------------------------------------------
// CPU-side self test of the FFT index algebra (no CUDA calls): DST-I of size n-1 computed by
// the very same pass / split functions the kernels use, threads emulated sequentially.
// ------------------------------------------------------------------------------------------
extern "C" int somax_b200_host_dst1_check(int n, const double* x /* n-1 */, double* X /* n-1 */) {
  using namespace sb;
  if (n < 8 || (n & (n - 1))) return -1;
  FftPlan plan = make_fft_plan(n);
  std::vector<C2<double>> tw(2 * (size_t)n), s(fft_padded_len(n));
  for (int t = 0; t < 2 * n; ++t) { tw[t].x = cos(-M_PI * t / n); tw[t].y = sin(-M_PI * t / n); }
  double* z = reinterpret_cast<double*>(s.data());
  auto zi = [&](int t) { return 2 * fft_pad(t >> 1) + (t & 1); };
  z[zi(0)] = 0; z[zi(n)] = 0;
  for (int t = 1; t < n; ++t) { z[zi(t)] = x[t - 1]; z[zi(2 * n - t)] = -x[t - 1]; }
  const int G = std::max(1, n / 8);
  int lgLc = plan.lgn;
  for (int ps = 0; ps < plan.npass; ++ps) {
    const int lr = plan.lgr[ps];
    for (int lt = 0; lt < G; ++lt) {
      if (lr == 3) fft_dif_pass<double, 8>(s.data(), plan.lgn, lgLc, lt, G, tw.data());
      else if (lr == 2) fft_dif_pass<double, 4>(s.data(), plan.lgn, lgLc, lt, G, tw.data());
      else fft_dif_pass<double, 2>(s.data(), plan.lgn, lgLc, lt, G, tw.data());
    }
    lgLc -= lr;
  }
  for (int k = 1; k <= n / 2; ++k) {
    double a, b;
    dst_split<double>(s.data(), plan, k, tw.data(), a, b);
    X[k - 1] = a;
    if (k != n - k) X[n - k - 1] = b;
  }
  return 0;
}
