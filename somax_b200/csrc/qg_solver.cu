// PV inversion for the QG models: psi = ring0(Cm2l . Helm^-1_lambda . Cl2m . q).
//
// Replaces BaroclinicQG._invert_pv / finitevolx.pv_inversion(bc="dst") (reference
// qg/baroclinic.py:135-159; recipe SURVEY.md App. B.4).  The reference transforms with a
// DST-I in both directions and divides by the 5-point eigenvalues, i.e. it solves the
// discrete system (delta_xx + delta_yy - lambda) psi = q with psi = 0 on the ghost ring
// EXACTLY.  Any exact direct solver of that system is therefore equivalent to rounding.
// DST-I of length nx needs an FFT of length 2(nx+1), which for nx = 2^p has a large prime
// factor (8193 = 3*2731), so this solver is B200-first instead:
//
//   FFT path (nx = 2^p): columns 1..nx-1 are transformed with a DST-I of size nx-1 (an
//     in-shared-memory radix-8 complex FFT of length nx); the last column is a border
//     handled by a Schur complement: solve 1 gives v, g = S^-1(f_n - b v(n-1)) via a dense
//     DST in y on that single column, solve 2 adds the harmonic correction for g.
//   dense path (any nx <= 2048): x transform as a dense DST-I matrix product.
//   Both: the y direction is solved per x-wavenumber by the Thomas algorithm, marching in y
//     with one thread per wavenumber (coalesced), carried in fp64 even for fp32 data (the
//     second-difference recurrences lose ~ (ny/pi)^2 eps otherwise), coefficients from a
//     compact host-built fp64 table (the Toeplitz recurrence converges to a fixed point).
#include "qg_solver.cuh"

#include <math.h>

#include <algorithm>
#include <vector>

#include "fft.cuh"

namespace sb {

constexpr int QG_MAX_NL = 4;
constexpr int TH_COLS = 64;   // columns (threads) per Thomas CTA
// cp.async ring depth (rows in flight per column): 16 KB of shared memory per CTA either way
template <typename T> struct ThDepth { static constexpr int v = sizeof(T) == 4 ? 64 : 32; };

struct Mix { double c[QG_MAX_NL][QG_MAX_NL]; };

struct QgSolver {
  int dtype, batch, nl, ny, nx, kind;
  double dx, dy;
  Layout L;
  int np, ncols, planes;
  void* S = nullptr; void* W = nullptr;
  double* ctab = nullptr; int* coff = nullptr; int* krow = nullptr; double* cinf = nullptr;
  int kbad[QG_MAX_NL] = {0, 0, 0, 0}; int KB = 0; double* dbad = nullptr;
  double* bsig = nullptr; double* sig2n = nullptr; double* sdiag = nullptr; double* sintab = nullptr;
  double* rvec = nullptr; double* ghat = nullptr; double* gvec = nullptr;
  void* tw = nullptr; void* dstmat = nullptr;
  FftPlan plan;
  Mix l2m, m2l;
  size_t bytes = 0;
};

// ------------------------------------------------------------------------------------------
// row kernels, FFT path
// ------------------------------------------------------------------------------------------
template <typename T>
struct RowArgs {
  Layout L;          // padded field layout
  int ny, n, np, nl; // n = nx
  int G, rows_per_block;
  FftPlan plan;
  T mix[QG_MAX_NL][QG_MAX_NL];
  const C2<T>* tw;
  T scale;
};

// forward: q (padded field) -> S[plane][j][0..n-2] = DST-I(n-1) of mode rows, S[..][n-1] = raw
// border column of the mode.
template <typename T>
__global__ void rowdst_fwd_fft(RowArgs<T> A, const T* __restrict__ q, T* __restrict__ S) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = A.n, G = A.G;
  const int lrow = threadIdx.x / G, lt = threadIdx.x - lrow * G;
  const int plen = fft_padded_len(n);
  C2<T>* s = reinterpret_cast<C2<T>*>(smem_raw) + (size_t)lrow * plen;
  T* z = reinterpret_cast<T*>(s);
  const int row = blockIdx.x * A.rows_per_block + lrow;   // (b, j) flattened
  const bool valid = row < A.L.batch * A.ny;
  const int b = valid ? row / A.ny : 0, j = valid ? row - b * A.ny : 0;
  auto zi = [&](int t) { return 2 * fft_pad(t >> 1) + (t & 1); };

  for (int m = 0; m < A.nl; ++m) {
    if (valid) {
      for (int p = lt; p < n; p += G) {           // p: interior index, x_{p+1}
        T val = 0;
        for (int l = 0; l < A.nl; ++l)
          val += A.mix[m][l] *
                 q[(((size_t)b * A.nl + l) * A.L.Ny + (j + 1)) * A.L.pitch + OFF + 1 + p];
        const int t = p + 1;
        if (t < n) { z[zi(t)] = val; z[zi(2 * n - t)] = -val; }
        else S[(((size_t)b * A.nl + m) * A.ny + j) * A.np + (n - 1)] = val;   // border column
      }
      if (lt == 0) { z[zi(0)] = 0; z[zi(n)] = 0; }
    }
    __syncthreads();
    int Lc = n;
    for (int ps = 0; ps < A.plan.npass; ++ps) {
      const int R = A.plan.radix[ps];
      if (valid) {
        if (R == 8) fft_dif_pass<T, 8>(s, n, Lc, lt, G, A.tw);
        else if (R == 4) fft_dif_pass<T, 4>(s, n, Lc, lt, G, A.tw);
        else fft_dif_pass<T, 2>(s, n, Lc, lt, G, A.tw);
      }
      Lc /= R;
      __syncthreads();
    }
    if (valid) {
      T* out = S + (((size_t)b * A.nl + m) * A.ny + j) * A.np;
      for (int k = lt + 1; k <= n / 2; k += G) {
        T Xk, Xnk;
        dst_split<T>(s, A.plan, k, A.tw, Xk, Xnk);
        out[k - 1] = Xk;
        if (k != n - k) out[n - k - 1] = Xnk;
      }
    }
    __syncthreads();
  }
}

// inverse: layer rows = (2/n) DST-I(n-1)[ sum_m Cm2l[l][m] U_m ], border column from slot n-1.
template <typename T>
__global__ void rowdst_inv_fft(RowArgs<T> A, const T* __restrict__ S, T* __restrict__ psi) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = A.n, G = A.G;
  const int lrow = threadIdx.x / G, lt = threadIdx.x - lrow * G;
  const int plen = fft_padded_len(n);
  C2<T>* s = reinterpret_cast<C2<T>*>(smem_raw) + (size_t)lrow * plen;
  T* z = reinterpret_cast<T*>(s);
  const int row = blockIdx.x * A.rows_per_block + lrow;
  const bool valid = row < A.L.batch * A.ny;
  const int b = valid ? row / A.ny : 0, j = valid ? row - b * A.ny : 0;
  auto zi = [&](int t) { return 2 * fft_pad(t >> 1) + (t & 1); };

  for (int l = 0; l < A.nl; ++l) {
    T* out = psi + (((size_t)b * A.nl + l) * A.L.Ny + (j + 1)) * A.L.pitch + OFF + 1;
    if (valid) {
      for (int p = lt; p < n; p += G) {           // p = k-1 for k = 1..n-1 ; p = n-1 border
        T val = 0;
        for (int m = 0; m < A.nl; ++m)
          val += A.mix[l][m] * S[(((size_t)b * A.nl + m) * A.ny + j) * A.np + p];
        const int t = p + 1;
        if (t < n) { z[zi(t)] = val; z[zi(2 * n - t)] = -val; }
        else out[n - 1] = val;                    // psi at the border column i = n
      }
      if (lt == 0) { z[zi(0)] = 0; z[zi(n)] = 0; }
    }
    __syncthreads();
    int Lc = n;
    for (int ps = 0; ps < A.plan.npass; ++ps) {
      const int R = A.plan.radix[ps];
      if (valid) {
        if (R == 8) fft_dif_pass<T, 8>(s, n, Lc, lt, G, A.tw);
        else if (R == 4) fft_dif_pass<T, 4>(s, n, Lc, lt, G, A.tw);
        else fft_dif_pass<T, 2>(s, n, Lc, lt, G, A.tw);
      }
      Lc /= R;
      __syncthreads();
    }
    if (valid) {
      for (int k = lt + 1; k <= n / 2; k += G) {
        T Xk, Xnk;
        dst_split<T>(s, A.plan, k, A.tw, Xk, Xnk);
        out[k - 1] = A.scale * Xk;
        if (k != n - k) out[n - k - 1] = A.scale * Xnk;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// row kernels, dense path: out[k] = scale * sum_i Smat[i*n + k] * (mixed row)[i]
// ------------------------------------------------------------------------------------------
template <typename T, bool INV>
__global__ void rowdst_dense(Layout L, int ny, int n, int np, int nl, Mix mix,
                             const T* __restrict__ Smat, const T* __restrict__ in,
                             T* __restrict__ out, double scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* xs = reinterpret_cast<T*>(smem_raw);   // nl * n
  const int row = blockIdx.x;
  const int b = row / ny, j = row - b * ny;
  for (int e = threadIdx.x; e < nl * n; e += blockDim.x) {
    int a = e / n, i = e - a * n;
    double val = 0;
    for (int c = 0; c < nl; ++c) {
      T src = INV ? in[(((size_t)b * nl + c) * ny + j) * np + i]
                  : in[(((size_t)b * nl + c) * L.Ny + (j + 1)) * L.pitch + OFF + 1 + i];
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
      if (INV) out[(((size_t)b * nl + a) * L.Ny + (j + 1)) * L.pitch + OFF + 1 + k] = (T)(scale * acc[a]);
      else out[(((size_t)b * nl + a) * ny + j) * np + k] = (T)(scale * acc[a]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Thomas sweeps along y, one thread per x-wavenumber, fp64 carry.
//   normalised system per column: x_{j-1} + delta x_j + x_{j+1} = dy^2 f_j
//   c_j = 1/(delta - c_{j-1}),  d_j = (dy^2 f_j - d_{j-1}) c_j,  x_j = d_j - c_j x_{j+1}
// ------------------------------------------------------------------------------------------
struct ThomasTab {
  const double* ctab; const int* coff; const int* krow; const double* cinf;
  int kbad[QG_MAX_NL]; int KB; double* dbad;
  int ny, np, ncols, nl;
  double dy2;
};

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem, const T* gmem) {
  if (sizeof(T) == 4) cp_async4(smem, gmem); else cp_async8(smem, gmem);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// FROM_VEC: right-hand side is gvec[plane][j] for every column (second, border solve).
template <typename T, bool FROM_VEC>
__global__ void __launch_bounds__(TH_COLS)
thomas_fwd(ThomasTab tb, const T* __restrict__ in, const double* __restrict__ gvec,
           T* __restrict__ out) {
  constexpr int TH_DEPTH = ThDepth<T>::v;
  __shared__ T ring[TH_DEPTH][TH_COLS];
  const int c = blockIdx.x * TH_COLS + threadIdx.x;
  const int plane = blockIdx.y, m = plane % tb.nl;
  const bool act = c < tb.ncols;
  const int cc = act ? c : 0;
  const size_t base = (size_t)plane * tb.ny * tb.np + cc;
  const double cfix = tb.cinf[m * tb.ncols + cc];
  const bool bad = act && cc < tb.kbad[m];
  if (!FROM_VEC) {
#pragma unroll 1
    for (int r = 0; r < TH_DEPTH; ++r) {
      if (r < tb.ny) cp_async_elem(&ring[r][threadIdx.x], in + base + (size_t)r * tb.np);
      cp_async_commit();
    }
  }
  double d = 0.0;
#pragma unroll 1
  for (int j = 0; j < tb.ny; ++j) {
    double f;
    if (FROM_VEC) {
      f = gvec[(size_t)plane * tb.ny + j];
    } else {
      cp_async_wait<TH_DEPTH - 1>();
      f = (double)ring[j % TH_DEPTH][threadIdx.x];
      if (j + TH_DEPTH < tb.ny)
        cp_async_elem(&ring[j % TH_DEPTH][threadIdx.x], in + base + (size_t)(j + TH_DEPTH) * tb.np);
      cp_async_commit();
    }
    const int kr = tb.krow[m * tb.ny + j];
    const double cj = (cc < kr) ? tb.ctab[(size_t)tb.coff[m * tb.ny + j] + cc] : cfix;
    d = (tb.dy2 * f - d) * cj;
    if (act) {
      out[base + (size_t)j * tb.np] = (T)d;
      if (bad) tb.dbad[((size_t)plane * tb.ny + j) * tb.KB + cc] = d;
    }
  }
}

// COMBINE: out = V - bsig[c] * x   (second solve applied to the first solve's result V)
template <typename T, bool COMBINE>
__global__ void __launch_bounds__(TH_COLS)
thomas_bwd(ThomasTab tb, const T* __restrict__ din, const T* __restrict__ V,
           const double* __restrict__ bsig, T* __restrict__ out) {
  constexpr int TH_DEPTH = ThDepth<T>::v;
  __shared__ T ring[TH_DEPTH][TH_COLS];
  __shared__ T ringv[COMBINE ? TH_DEPTH : 1][TH_COLS];
  const int c = blockIdx.x * TH_COLS + threadIdx.x;
  const int plane = blockIdx.y, m = plane % tb.nl;
  const bool act = c < tb.ncols;
  const int cc = act ? c : 0;
  const size_t base = (size_t)plane * tb.ny * tb.np + cc;
  const double cfix = tb.cinf[m * tb.ncols + cc];
  const bool bad = act && cc < tb.kbad[m];
  const double bs = COMBINE ? bsig[cc] : 0.0;
#pragma unroll 1
  for (int r = 0; r < TH_DEPTH; ++r) {
    const int j = tb.ny - 1 - r;
    if (j >= 0) {
      cp_async_elem(&ring[r][threadIdx.x], din + base + (size_t)j * tb.np);
      if (COMBINE) cp_async_elem(&ringv[r][threadIdx.x], V + base + (size_t)j * tb.np);
    }
    cp_async_commit();
  }
  double x = 0.0;
#pragma unroll 1
  for (int r = 0; r < tb.ny; ++r) {
    const int j = tb.ny - 1 - r;
    cp_async_wait<TH_DEPTH - 1>();
    double d = (double)ring[r % TH_DEPTH][threadIdx.x];
    double v = COMBINE ? (double)ringv[r % TH_DEPTH][threadIdx.x] : 0.0;
    const int jn = j - TH_DEPTH;
    if (jn >= 0) {
      cp_async_elem(&ring[r % TH_DEPTH][threadIdx.x], din + base + (size_t)jn * tb.np);
      if (COMBINE) cp_async_elem(&ringv[r % TH_DEPTH][threadIdx.x], V + base + (size_t)jn * tb.np);
    }
    cp_async_commit();
    if (bad) d = tb.dbad[((size_t)plane * tb.ny + j) * tb.KB + cc];
    const int kr = tb.krow[m * tb.ny + j];
    const double cj = (cc < kr) ? tb.ctab[(size_t)tb.coff[m * tb.ny + j] + cc] : cfix;
    x = d - cj * x;                     // x_{ny+1} = 0
    if (act) out[base + (size_t)j * tb.np] = COMBINE ? (T)(v - bs * x) : (T)x;
  }
}

// r[plane][j] = sum_c sig2n[c] * V[plane][j][c]   (= value of the first solve at column n-1)
template <typename T>
__global__ void border_dot(const T* __restrict__ V, const double* __restrict__ sig2n, int ny,
                           int np, int ncols, double* __restrict__ r) {
  const int j = blockIdx.x, plane = blockIdx.y;
  const T* row = V + ((size_t)plane * ny + j) * np;
  double acc = 0;
  for (int c = threadIdx.x; c < ncols; c += blockDim.x) acc += sig2n[c] * (double)row[c];
  __shared__ double red[32];
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    r[(size_t)plane * ny + j] = t;
  }
}

// Border Schur solve, stage A: ghat[l] = (sum_j sin(pi j l/N) (f_n[j] - b r[j])) / sdiag[m][l]
template <typename T>
__global__ void border_gsolve_a(const T* __restrict__ S, const double* __restrict__ r,
                                const double* __restrict__ sintab, const double* __restrict__ sdiag,
                                int ny, int np, int n, int nl, double b, double* __restrict__ ghat) {
  const int l = blockIdx.x + 1, plane = blockIdx.y, m = plane % nl;
  const int N2 = 2 * (ny + 1);
  double acc = 0;
  for (int j = threadIdx.x + 1; j <= ny; j += blockDim.x) {
    double rhs = (double)S[((size_t)plane * ny + (j - 1)) * np + (n - 1)] - b * r[(size_t)plane * ny + (j - 1)];
    int idx = (int)(((long long)j * l) % N2);
    acc += sintab[idx] * rhs;
  }
  __shared__ double red[32];
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    ghat[(size_t)plane * ny + (l - 1)] = t / sdiag[(size_t)m * ny + (l - 1)];
  }
}

// stage B: g[j] = (2/N) sum_l sin(pi j l/N) ghat[l]; stored to gvec (fp64) and the border slot.
template <typename T>
__global__ void border_gsolve_b(const double* __restrict__ ghat, const double* __restrict__ sintab,
                                int ny, int np, int n, double* __restrict__ gvec, T* __restrict__ S) {
  const int j = blockIdx.x + 1, plane = blockIdx.y;
  const int N2 = 2 * (ny + 1);
  double acc = 0;
  for (int l = threadIdx.x + 1; l <= ny; l += blockDim.x) {
    int idx = (int)(((long long)j * l) % N2);
    acc += sintab[idx] * ghat[(size_t)plane * ny + (l - 1)];
  }
  __shared__ double red[32];
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    t *= 2.0 / (ny + 1);
    gvec[(size_t)plane * ny + (j - 1)] = t;
    S[((size_t)plane * ny + (j - 1)) * np + (n - 1)] = (T)t;
  }
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
  const int nl = s->nl, ny = s->ny, nc = s->ncols;
  std::vector<double> cinf((size_t)nl * nc);
  std::vector<int> J((size_t)nl * nc);       // rows 1..J use the table
  std::vector<std::vector<double>> cols((size_t)nl * nc);
  const double dy2 = s->dy * s->dy;
  s->KB = 0;
  for (int m = 0; m < nl; ++m) {
    s->kbad[m] = 0;
    for (int c = 0; c < nc; ++c) {
      const double sn = sin(M_PI * (c + 1) / (2.0 * Nx_eig));
      const double lam_x = -(4.0 / (s->dx * s->dx)) * sn * sn;
      const double delta = (lam_x - lambdas[m]) * dy2 - 2.0;
      auto& col = cols[(size_t)m * nc + c];
      const bool definite = delta < -2.0;
      double cstar = 0.0;
      if (definite) cstar = 0.5 * (delta + sqrt(delta * delta - 4.0));
      else s->kbad[m] = std::max(s->kbad[m], c + 1);
      cinf[(size_t)m * nc + c] = cstar;
      double cj = 0.0;
      int jconv = ny;
      for (int j = 1; j <= ny; ++j) {
        cj = 1.0 / (delta - cj);
        if (definite && fabs(cj - cstar) <= 4e-16 * fabs(cstar)) { jconv = j - 1; break; }
        col.push_back(cj);
      }
      J[(size_t)m * nc + c] = jconv;
    }
    s->KB = std::max(s->KB, s->kbad[m]);
  }
  // row-major ragged table: row j (1-based) of mode m holds columns 0..K-1, K = 1 + max{c: J>=j}
  std::vector<int> krow((size_t)nl * ny), coff((size_t)nl * ny);
  std::vector<double> ctab;
  for (int m = 0; m < nl; ++m) {
    // K_j is non-increasing in j; compute via suffix sweep
    std::vector<int> K(ny + 2, 0);
    for (int c = 0; c < nc; ++c) {
      int jc = J[(size_t)m * nc + c];
      if (jc >= 1) K[jc] = std::max(K[jc], c + 1);
    }
    for (int j = ny - 1; j >= 1; --j) K[j] = std::max(K[j], K[j + 1]);
    for (int j = 1; j <= ny; ++j) {
      krow[(size_t)m * ny + (j - 1)] = K[j];
      coff[(size_t)m * ny + (j - 1)] = (int)ctab.size();
      for (int c = 0; c < K[j]; ++c) {
        const auto& col = cols[(size_t)m * nc + c];
        ctab.push_back(j <= (int)col.size() ? col[j - 1] : cinf[(size_t)m * nc + c]);
      }
    }
  }
  if (ctab.size() > (size_t)2000000000) return fail(SOMAX_B200_ERR_UNSUPPORTED, "thomas table too large");
  if (int rc = dev_upload(ctab.data(), ctab.size() * 8, (void**)&s->ctab, &s->bytes)) return rc;
  if (int rc = dev_upload(krow.data(), krow.size() * 4, (void**)&s->krow, &s->bytes)) return rc;
  if (int rc = dev_upload(coff.data(), coff.size() * 4, (void**)&s->coff, &s->bytes)) return rc;
  if (int rc = dev_upload(cinf.data(), cinf.size() * 8, (void**)&s->cinf, &s->bytes)) return rc;
  if (s->KB > 0) {
    size_t nb = (size_t)s->planes * ny * s->KB * 8;
    SB_CUDA(cudaMalloc((void**)&s->dbad, nb));
    SB_CUDA(cudaMemset(s->dbad, 0, nb));
    s->bytes += nb;
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
  std::vector<double> sig(nc), lamx(nc), bsig(nc), sig2n(nc);
  for (int c = 0; c < nc; ++c) {
    const int k = c + 1;
    sig[c] = ((k & 1) ? 1.0 : -1.0) * sin(M_PI * k / n);
    const double sn = sin(M_PI * k / (2.0 * n));
    lamx[c] = -(4.0 * b) * sn * sn;
    bsig[c] = b * sig[c];
    sig2n[c] = (2.0 / n) * sig[c];
  }
  std::vector<double> sdiag((size_t)nl * ny);
  for (int m = 0; m < nl; ++m)
    for (int l = 1; l <= ny; ++l) {
      const double sn = sin(M_PI * l / (2.0 * (ny + 1)));
      const double mu = -(4.0 / (s->dy * s->dy)) * sn * sn - lambdas[m];
      double acc = 0;
      for (int c = 0; c < nc; ++c) acc += sig[c] * sig[c] / (lamx[c] + mu);
      sdiag[(size_t)m * ny + (l - 1)] = mu - 2.0 * b - b * b * (2.0 / n) * acc;
    }
  std::vector<double> sintab(2 * (size_t)(ny + 1));
  for (size_t t = 0; t < sintab.size(); ++t) sintab[t] = sin(M_PI * (double)t / (ny + 1));
  if (int rc = dev_upload(bsig.data(), nc * 8, (void**)&s->bsig, &s->bytes)) return rc;
  if (int rc = dev_upload(sig2n.data(), nc * 8, (void**)&s->sig2n, &s->bytes)) return rc;
  if (int rc = dev_upload(sdiag.data(), sdiag.size() * 8, (void**)&s->sdiag, &s->bytes)) return rc;
  if (int rc = dev_upload(sintab.data(), sintab.size() * 8, (void**)&s->sintab, &s->bytes)) return rc;
  size_t vb = (size_t)s->planes * ny * 8;
  SB_CUDA(cudaMalloc((void**)&s->rvec, vb));
  SB_CUDA(cudaMalloc((void**)&s->ghat, vb));
  SB_CUDA(cudaMalloc((void**)&s->gvec, vb));
  s->bytes += 3 * vb;
  return 0;
}

template <typename T>
static int build_dense_tables(QgSolver* s) {
  const int n = s->nx;
  std::vector<T> mat((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k)
      mat[(size_t)i * n + k] = (T)sin(M_PI * (double)(i + 1) * (double)(k + 1) / (n + 1));
  return dev_upload(mat.data(), mat.size() * sizeof(T), &s->dstmat, &s->bytes);
}

int qg_solver_create(QgSolver** out, int dtype, int batch, int nl, int ny, int nx, double dx,
                     double dy, const double* Cl2m, const double* Cm2l, const double* lambdas,
                     int solver_kind) {
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
    if (nx > 2048) return fail(SOMAX_B200_ERR_UNSUPPORTED, "dense DST solver limited to nx <= 2048; use nx = 2^p for the FFT path");
  } else {
    return fail(SOMAX_B200_ERR_INVALID, "unknown solver kind");
  }
  auto* s = new QgSolver();
  s->dtype = dtype; s->batch = batch; s->nl = nl; s->ny = ny; s->nx = nx; s->kind = kind;
  s->dx = dx; s->dy = dy; s->L = make_layout(batch, nl, ny, nx);
  s->np = ((nx + 3) / 4) * 4; s->planes = batch * nl;
  s->ncols = (kind == SOMAX_B200_SOLVER_FFT) ? nx - 1 : nx;
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
  if (!rc) rc = build_thomas_tables(s, lambdas, kind == SOMAX_B200_SOLVER_FFT ? nx : nx + 1);
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
  void* ptrs[] = {s->S, s->W, s->ctab, s->coff, s->krow, s->cinf, s->dbad, s->bsig, s->sig2n,
                  s->sdiag, s->sintab, s->rvec, s->ghat, s->gvec, s->tw, s->dstmat};
  for (void* p : ptrs) cudaFree(p);
  delete s;
}

size_t qg_solver_bytes(const QgSolver* s) { return s ? s->bytes : 0; }
int qg_solver_kind(const QgSolver* s) { return s->kind; }

template <typename T>
static RowArgs<T> make_row_args(QgSolver* s, const Mix& mix, double scale) {
  RowArgs<T> A;
  A.L = s->L; A.ny = s->ny; A.n = s->nx; A.np = s->np; A.nl = s->nl; A.plan = s->plan;
  const int n = s->nx;
  int ept = n >= 8192 ? 16 : (n >= 256 ? 8 : (n >= 128 ? 4 : 2));
  A.G = n / ept;
  int threads = std::max(A.G, 128);
  A.rows_per_block = threads / A.G;
  // keep shared memory of one block within the opt-in limit
  const size_t row_bytes = (size_t)fft_padded_len(n) * sizeof(C2<T>);
  while (A.rows_per_block > 1 && row_bytes * A.rows_per_block > 96 * 1024) A.rows_per_block /= 2;
  for (int a = 0; a < QG_MAX_NL; ++a)
    for (int c = 0; c < QG_MAX_NL; ++c) A.mix[a][c] = (T)mix.c[a][c];
  A.tw = (const C2<T>*)s->tw;
  A.scale = (T)scale;
  return A;
}

template <typename T>
int qg_solver_run(QgSolver* s, const T* q, T* psi, cudaStream_t st) {
  const int ny = s->ny, n = s->nx, np = s->np, nl = s->nl;
  ThomasTab tb;
  tb.ctab = s->ctab; tb.coff = s->coff; tb.krow = s->krow; tb.cinf = s->cinf;
  for (int m = 0; m < QG_MAX_NL; ++m) tb.kbad[m] = s->kbad[m];
  tb.KB = s->KB; tb.dbad = s->dbad; tb.ny = ny; tb.np = np; tb.ncols = s->ncols; tb.nl = nl;
  tb.dy2 = s->dy * s->dy;
  dim3 tgrid((s->ncols + TH_COLS - 1) / TH_COLS, s->planes);
  T* S = (T*)s->S;
  if (s->kind == SOMAX_B200_SOLVER_FFT) {
    T* W = (T*)s->W;
    RowArgs<T> Af = make_row_args<T>(s, s->l2m, 1.0);
    RowArgs<T> Ai = make_row_args<T>(s, s->m2l, 2.0 / n);
    const size_t smem = (size_t)fft_padded_len(n) * sizeof(C2<T>) * Af.rows_per_block;
    const int threads = Af.G * Af.rows_per_block;
    const int nrows = s->batch * ny;
    const int blocks = (nrows + Af.rows_per_block - 1) / Af.rows_per_block;
    static thread_local size_t smem_set_f = 0, smem_set_i = 0;
    if (smem > 48 * 1024) {
      if (smem_set_f < smem) {
        SB_CUDA(cudaFuncSetAttribute(rowdst_fwd_fft<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set_f = smem;
      }
      if (smem_set_i < smem) {
        SB_CUDA(cudaFuncSetAttribute(rowdst_inv_fft<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set_i = smem;
      }
    }
    prof_begin("rowdst_fwd_fft", st);
    rowdst_fwd_fft<T><<<blocks, threads, smem, st>>>(Af, q, S);
    SB_LAUNCH_CHECK();
    prof_begin("thomas_fwd_0", st);
    thomas_fwd<T, false><<<tgrid, TH_COLS, 0, st>>>(tb, S, nullptr, S);
    SB_LAUNCH_CHECK();
    prof_begin("thomas_bwd_0", st);
    thomas_bwd<T, false><<<tgrid, TH_COLS, 0, st>>>(tb, S, nullptr, nullptr, S);
    SB_LAUNCH_CHECK();
    prof_begin("border_dot", st);
    border_dot<T><<<dim3(ny, s->planes), 256, 0, st>>>(S, s->sig2n, ny, np, s->ncols, s->rvec);
    SB_LAUNCH_CHECK();
    const double b = 1.0 / (s->dx * s->dx);
    prof_begin("border_gsolve_a", st);
    border_gsolve_a<T><<<dim3(ny, s->planes), 128, 0, st>>>(S, s->rvec, s->sintab, s->sdiag, ny, np, n, nl, b, s->ghat);
    SB_LAUNCH_CHECK();
    prof_begin("border_gsolve_b", st);
    border_gsolve_b<T><<<dim3(ny, s->planes), 128, 0, st>>>(s->ghat, s->sintab, ny, np, n, s->gvec, S);
    SB_LAUNCH_CHECK();
    prof_begin("thomas_fwd_1", st);
    thomas_fwd<T, true><<<tgrid, TH_COLS, 0, st>>>(tb, nullptr, s->gvec, W);
    SB_LAUNCH_CHECK();
    prof_begin("thomas_bwd_1", st);
    thomas_bwd<T, true><<<tgrid, TH_COLS, 0, st>>>(tb, W, S, s->bsig, S);
    SB_LAUNCH_CHECK();
    prof_begin("rowdst_inv_fft", st);
    rowdst_inv_fft<T><<<blocks, threads, smem, st>>>(Ai, S, psi);
    SB_LAUNCH_CHECK();
  } else {
    const size_t smem = (size_t)nl * n * sizeof(T);
    if (smem > 48 * 1024) {
      SB_CUDA(cudaFuncSetAttribute(rowdst_dense<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      SB_CUDA(cudaFuncSetAttribute(rowdst_dense<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int threads = std::min(256, ((n + 31) / 32) * 32);
    prof_begin("rowdst_dense_0", st);
    rowdst_dense<T, false><<<s->batch * ny, threads, smem, st>>>(s->L, ny, n, np, nl, s->l2m, (const T*)s->dstmat, q, S, 1.0);
    SB_LAUNCH_CHECK();
    prof_begin("thomas_fwd_0", st);
    thomas_fwd<T, false><<<tgrid, TH_COLS, 0, st>>>(tb, S, nullptr, S);
    SB_LAUNCH_CHECK();
    prof_begin("thomas_bwd_0", st);
    thomas_bwd<T, false><<<tgrid, TH_COLS, 0, st>>>(tb, S, nullptr, nullptr, S);
    SB_LAUNCH_CHECK();
    prof_begin("rowdst_dense_1", st);
    rowdst_dense<T, true><<<s->batch * ny, threads, smem, st>>>(s->L, ny, n, np, nl, s->m2l, (const T*)s->dstmat, S, psi, 2.0 / (n + 1));
    SB_LAUNCH_CHECK();
  }
  return 0;
}

template int qg_solver_run<float>(QgSolver*, const float*, float*, cudaStream_t);
template int qg_solver_run<double>(QgSolver*, const double*, double*, cudaStream_t);

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
  int Lc = n;
  for (int ps = 0; ps < plan.npass; ++ps) {
    const int R = plan.radix[ps];
    for (int lt = 0; lt < G; ++lt) {
      if (R == 8) fft_dif_pass<double, 8>(s.data(), n, Lc, lt, G, tw.data());
      else if (R == 4) fft_dif_pass<double, 4>(s.data(), n, Lc, lt, G, tw.data());
      else fft_dif_pass<double, 2>(s.data(), n, Lc, lt, G, tw.data());
    }
    Lc /= R;
  }
  for (int k = 1; k <= n / 2; ++k) {
    double a, b;
    dst_split<double>(s.data(), plan, k, tw.data(), a, b);
    X[k - 1] = a;
    if (k != n - k) X[n - k - 1] = b;
  }
  return 0;
}
