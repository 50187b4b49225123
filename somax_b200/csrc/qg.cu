// Quasi-geostrophic hot path: fused Arakawa-Jacobian / beta / viscosity / wind / drag stencil
// with the Tsit5 stage combination in the epilogue, driven around the PV-inversion solver.
//
// Replaces BaroclinicQG.{apply_boundary_conditions,_invert_pv,vector_field,diagnose} and
// BarotropicQG's equivalents plus the diffrax Tsit5 loop (reference qg/baroclinic.py:135-228,
// qg/barotropic.py:113-183, core/model.py:47-88; Jacobian formula SURVEY.md App. B.3).
#include "common.cuh"
#include "qg_solver.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

namespace sb {

constexpr int QTY = 8;             // output rows per CTA
constexpr int QTXG = 32;           // float4 groups per CTA row
constexpr int QTW = QTXG * 4 + 2;  // tile width with halo
constexpr int QTWP = QTW + 2;

template <typename T>
struct QgArgs {
  Layout L;
  int apply_bc;
  int bc_ylo, bc_yhi;   // row 0 / row Ny-1 is the physical ring (0 on a slab whose neighbour owns it: halo row)
  T dx2, dy2, jden;   // dx^2, dy^2, 12 dx dy
  T idx2, idy2, ijden, iH0;   // reciprocals (fast kernel)
  const T* beta; int b_cp, b_xs;
  const T* wind; int w_cp, w_xs;
  T H0, nu, kappa, tau0;
  unsigned nl_magic;   // ceil(2^32 / nl): plane / nl = umulhi(plane, nl_magic) for plane < 2^16
};

template <typename T>
__global__ void __launch_bounds__(QTXG* QTY)
qg_rhs_kernel(QgArgs<T> A, const T* __restrict__ psi, Stage<T> st) {
  __shared__ T s_q[QTY + 2][QTWP];   // q (BC applied)
  __shared__ T s_t[QTY + 2][QTWP];   // q + beta_y
  __shared__ T s_p[QTY + 2][QTWP];   // psi
  const Layout& L = A.L;
  const int Ny = L.Ny, Nx = L.Nx, pitch = L.pitch;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * QTXG + tx;
  const int g0 = blockIdx.x * QTXG, c0 = g0 * 4 - OFF, j0 = blockIdx.y * QTY;
  const int plane = blockIdx.z, k = plane % L.nl;
  const size_t po = (size_t)plane * L.plane();
  const T* pq = st.Yin[0] + po;
  const T* pp = psi + po;

  for (int e = tid; e < (QTY + 2) * QTW; e += QTXG * QTY) {
    int r = e / QTW, c = e - r * QTW;
    int jj = j0 - 1 + r, ii = c0 - 1 + c;
    T q = 0, t = 0, p = 0;
    if (jj >= 0 && jj < Ny && ii >= 0 && ii < Nx) {
      const bool ring = ((jj == 0 && A.bc_ylo) || (jj == Ny - 1 && A.bc_yhi) || ii == 0 || ii == Nx - 1);
      size_t o = (size_t)jj * pitch + OFF + ii;
      q = (A.apply_bc && ring) ? T(0) : pq[o];
      t = q + A.beta[(size_t)jj * A.b_cp + (size_t)ii * A.b_xs];
      p = pp[o];
    }
    s_q[r][c] = q; s_t[r][c] = t; s_p[r][c] = p;
  }
  __syncthreads();
  const int j = j0 + ty, g = g0 + tx;
  if (j >= Ny || g >= L.groups()) return;
  const int r = ty + 1;
  T out[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = c0 + tx * 4 + e, c = tx * 4 + e + 1;
    T dq = 0;
    if (i >= 0 && i < Nx) {
      const bool interior = (j >= 1 && j <= Ny - 2 && i >= 1 && i <= Nx - 2);
      if (interior) {
        auto F = [&](int dr, int dc) { return s_p[r + dr][c + dc]; };
        auto G = [&](int dr, int dc) { return s_t[r + dr][c + dc]; };
        const T fE = F(0, 1), fW = F(0, -1), fN = F(1, 0), fS = F(-1, 0);
        const T fNE = F(1, 1), fNW = F(1, -1), fSE = F(-1, 1), fSW = F(-1, -1);
        const T gE = G(0, 1), gW = G(0, -1), gN = G(1, 0), gS = G(-1, 0);
        const T gNE = G(1, 1), gNW = G(1, -1), gSE = G(-1, 1), gSW = G(-1, -1);
        const T jpp = (fE - fW) * (gN - gS) - (fN - fS) * (gE - gW);
        const T jpx = fE * (gNE - gSE) - fW * (gNW - gSW) - fN * (gNE - gNW) + fS * (gSE - gSW);
        const T jxp = gN * (fNE - fNW) - gS * (fSE - fSW) - gE * (fNE - fSE) + gW * (fNW - fSW);
        dq = -(((jpp + jpx) + jxp) / A.jden);
      }
      if (k == 0) dq = dq + (A.tau0 * A.wind[(size_t)j * A.w_cp + (size_t)i * A.w_xs]) / A.H0;
      if (interior) {
        if (k == L.nl - 1) {
          const T pc = s_p[r][c];
          const T lap = (s_p[r][c + 1] - T(2) * pc + s_p[r][c - 1]) / A.dx2 +
                        (s_p[r + 1][c] - T(2) * pc + s_p[r - 1][c]) / A.dy2;
          dq = dq + (-A.kappa * lap);
        }
        const T qc = s_q[r][c];
        const T lapq = (s_q[r][c + 1] - T(2) * qc + s_q[r][c - 1]) / A.dx2 +
                       (s_q[r + 1][c] - T(2) * qc + s_q[r - 1][c]) / A.dy2;
        dq = dq + A.nu * lapq;
      }
    }
    out[e] = dq;
  }
  const size_t idx = po + (size_t)j * pitch + (size_t)g * 4;
  const bool need_yin = (st.Yout[0] != nullptr) && (st.y[0] == nullptr);
  Vec4<T> yin = need_yin ? ld4(st.Yin[0] + idx) : Vec4<T>{0, 0, 0, 0};
  rk_epilogue4(st, 0, idx, yin, Vec4<T>{out[0], out[1], out[2], out[3]});
}

// Fast variant used when beta_y and the wind pattern depend on y only (every reference factory):
// 128-bit global->shared tile loads, 128-bit shared reads, warp shuffles for the x neighbours.
constexpr int QSW = QTXG * 4 + 8;   // shared row: 4 slots left (halo at [3]) | 128 | 4 slots right (halo at [0])

template <typename T>
__global__ void __launch_bounds__(QTXG* QTY)
qg_rhs_kernel_fast(QgArgs<T> A, const T* __restrict__ psi, Stage<T> st) {
  __shared__ __align__(32) T s_q[QTY + 2][QSW];
  __shared__ __align__(32) T s_p[QTY + 2][QSW];
  const Layout& L = A.L;
  const int Ny = L.Ny, Nx = L.Nx, pitch = L.pitch, ngroups = L.groups();
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * QTXG + tx;
  const int g0 = blockIdx.x * QTXG, j0 = blockIdx.y * QTY;
  const int plane = blockIdx.z, k = plane % L.nl;
  const size_t po = (size_t)plane * L.plane();
  const T* pq = st.Yin[0] + po;
  const T* pp = psi + po;
  const size_t idx = po + (size_t)(j0 + ty) * pitch + (size_t)(g0 + tx) * 4;
  // fp32: the epilogue's operands (step start state and the previous stage derivatives) are
  // prefetched with 16-byte cp.async into per-thread shared slots right away, so their HBM
  // latency overlaps the tile load, the barrier and the stencil arithmetic without holding
  // registers.
  constexpr bool STAGE_EPI = sizeof(T) == 4;
  __shared__ __align__(16) T s_epi[STAGE_EPI ? MAX_PREV + 1 : 1][QTXG * QTY][STAGE_EPI ? 4 : 1];
  const bool epi_valid = (j0 + ty) < Ny && (g0 + tx) < ngroups && st.Yout[0] != nullptr;
  if (STAGE_EPI && epi_valid) {
    const T* base = (st.y[0] ? st.y[0] : st.Yin[0]) + idx;
    unsigned sa = (unsigned)__cvta_generic_to_shared(&s_epi[MAX_PREV][tid][0]);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(base) : "memory");
#pragma unroll
    for (int jj = 0; jj < MAX_PREV; ++jj)
      if (jj < st.nprev) {
        sa = (unsigned)__cvta_generic_to_shared(&s_epi[jj][tid][0]);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(st.Fprev[jj][0] + idx) : "memory");
      }
  }
  if (STAGE_EPI) asm volatile("cp.async.commit_group;\n" ::: "memory");

  for (int e = tid; e < (QTY + 2) * (QTXG + 2); e += QTXG * QTY) {
    const int r = e / (QTXG + 2), gs = e - r * (QTXG + 2);   // gs: shared group 0..QTXG+1
    const int jj = j0 - 1 + r, gg = g0 - 1 + gs;
    Vec4<T> q{0, 0, 0, 0}, p{0, 0, 0, 0};
    if (jj >= 0 && jj < Ny && gg >= 0 && gg < ngroups) {
      const size_t o = (size_t)jj * pitch + (size_t)gg * 4;
      q = ld4(pq + o);
      p = ld4(pp + o);
      if (A.apply_bc) {
        const int i0 = gg * 4 - OFF;     // column of .x
        if ((jj == 0 && A.bc_ylo) || (jj == Ny - 1 && A.bc_yhi)) q = Vec4<T>{0, 0, 0, 0};
        if (i0 == 0 || i0 == Nx - 1) q.x = 0;
        if (i0 + 1 == 0 || i0 + 1 == Nx - 1) q.y = 0;
        if (i0 + 2 == 0 || i0 + 2 == Nx - 1) q.z = 0;
        if (i0 + 3 == 0 || i0 + 3 == Nx - 1) q.w = 0;
      }
    }
    st4(&s_q[r][gs * 4], q);
    st4(&s_p[r][gs * 4], p);
  }
  __syncthreads();
  const int j = j0 + ty, g = g0 + tx;
  if (j >= Ny) return;       // warp-uniform (ty); lanes with g >= ngroups stay for the shuffles
  const int r = ty + 1, cs = 4 * (tx + 1);
  // 3 x 6 windows of psi (f) and q (g): columns cs-1 .. cs+4
  T f[3][6], q[3][6];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const Vec4<T> a = ld4(&s_p[r - 1 + d][cs]);
    const Vec4<T> b = ld4(&s_q[r - 1 + d][cs]);
    T al = __shfl_up_sync(0xffffffffu, a.w, 1), ar = __shfl_down_sync(0xffffffffu, a.x, 1);
    T bl = __shfl_up_sync(0xffffffffu, b.w, 1), br = __shfl_down_sync(0xffffffffu, b.x, 1);
    if (tx == 0) { al = s_p[r - 1 + d][cs - 1]; bl = s_q[r - 1 + d][cs - 1]; }
    if (tx == QTXG - 1) { ar = s_p[r - 1 + d][cs + 4]; br = s_q[r - 1 + d][cs + 4]; }
    f[d][0] = al; f[d][1] = a.x; f[d][2] = a.y; f[d][3] = a.z; f[d][4] = a.w; f[d][5] = ar;
    q[d][0] = bl; q[d][1] = b.x; q[d][2] = b.y; q[d][3] = b.z; q[d][4] = b.w; q[d][5] = br;
  }
  if (g >= ngroups) return;
  // beta_y(row) for rows j-1, j, j+1 (clamped rows are never used by interior cells)
  T bet[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    int jj = j - 1 + d;
    jj = jj < 0 ? 0 : (jj > Ny - 1 ? Ny - 1 : jj);
    bet[d] = A.beta[jj];
  }
  const T wind = (k == 0) ? (A.tau0 * A.wind[j]) * A.iH0 : T(0);
  const int i0 = g * 4 - OFF;
  T out[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = i0 + e, c = e + 1;
    T dq = 0;
    if (i >= 0 && i < Nx) {
      const bool interior = (j >= 1 && j <= Ny - 2 && i >= 1 && i <= Nx - 2);
      if (interior) {
        const T fE = f[1][c + 1], fW = f[1][c - 1], fN = f[2][c], fS = f[0][c];
        const T fNE = f[2][c + 1], fNW = f[2][c - 1], fSE = f[0][c + 1], fSW = f[0][c - 1];
        const T gE = q[1][c + 1] + bet[1], gW = q[1][c - 1] + bet[1];
        const T gN = q[2][c] + bet[2], gS = q[0][c] + bet[0];
        const T gNE = q[2][c + 1] + bet[2], gNW = q[2][c - 1] + bet[2];
        const T gSE = q[0][c + 1] + bet[0], gSW = q[0][c - 1] + bet[0];
        const T jpp = (fE - fW) * (gN - gS) - (fN - fS) * (gE - gW);
        const T jpx = fE * (gNE - gSE) - fW * (gNW - gSW) - fN * (gNE - gNW) + fS * (gSE - gSW);
        const T jxp = gN * (fNE - fNW) - gS * (fSE - fSW) - gE * (fNE - fSE) + gW * (fNW - fSW);
        dq = -(((jpp + jpx) + jxp) * A.ijden);
      }
      if (k == 0) dq = dq + wind;
      if (interior) {
        if (k == L.nl - 1) {
          const T pc = f[1][c];
          const T lap = (f[1][c + 1] - T(2) * pc + f[1][c - 1]) * A.idx2 +
                        (f[2][c] - T(2) * pc + f[0][c]) * A.idy2;
          dq = dq + (-A.kappa * lap);
        }
        const T qc = q[1][c];
        const T lapq = (q[1][c + 1] - T(2) * qc + q[1][c - 1]) * A.idx2 +
                       (q[2][c] - T(2) * qc + q[0][c]) * A.idy2;
        dq = dq + A.nu * lapq;
      }
    }
    out[e] = dq;
  }
  const Vec4<T> F{out[0], out[1], out[2], out[3]};
  if (!STAGE_EPI) { rk_epilogue4_fast(st, 0, idx, F); return; }
  if (st.Fout[0]) st4(st.Fout[0] + idx, F);
  if (!st.Yout[0]) return;
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  Vec4<T> acc = ld4(&s_epi[MAX_PREV][tid][0]);
#pragma unroll
  for (int jj = 0; jj < MAX_PREV; ++jj) {
    if (jj < st.nprev) {
      const Vec4<T> kk = ld4(&s_epi[jj][tid][0]);
      acc.x = fma(st.adt[jj], kk.x, acc.x); acc.y = fma(st.adt[jj], kk.y, acc.y);
      acc.z = fma(st.adt[jj], kk.z, acc.z); acc.w = fma(st.adt[jj], kk.w, acc.w);
    }
  }
  acc.x = fma(st.adt_new, F.x, acc.x); acc.y = fma(st.adt_new, F.y, acc.y);
  acc.z = fma(st.adt_new, F.z, acc.z); acc.w = fma(st.adt_new, F.w, acc.w);
  st4(st.Yout[0] + idx, acc);
}

// ------------------------------------------------------------------------------------------
// fp32 stencil with TMA tile loads (cp.async.bulk.tensor, one elected thread): the q and psi
// halo tiles and the Tsit5 epilogue operands (step-start state + previous stage derivatives)
// arrive in shared memory through 3-D tensor maps over the padded (pitch, Ny, plane) arrays --
// out-of-range boxes are zero-filled by the hardware, so the loader has no per-thread address
// arithmetic or bounds checks at all.  Each thread computes 2 rows x 4 columns from a 4 x 6
// register window; only tiles that reach past the last interior row / column take the predicated
// EDGE path (none on a power-of-two grid: the tile grid starts at the first interior cell).
// Same arithmetic (operation order) as qg_rhs_kernel_fast.
// ------------------------------------------------------------------------------------------
constexpr int QFR = 2;                    // rows per thread
constexpr int QFY = QTY * QFR;            // output rows per CTA
constexpr int QHW = (QTXG + 2) * 4;       // halo tile width (floats): one group either side
constexpr int QHR = QFY + 2;              // halo tile rows
constexpr size_t QG_TMA_HALO_B = ((size_t)QHR * QHW * 4 + 127) / 128 * 128;
constexpr size_t QG_TMA_EPI_B = (size_t)QFY * QTXG * 4 * 4;
constexpr size_t QG_TMA_SMEM = 2 * QG_TMA_HALO_B + (MAX_PREV + 1) * QG_TMA_EPI_B + 16;

struct QgTmaps { CUtensorMap q, psi, base, f[MAX_PREV]; };

__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* tm, int x, int y, int z,
                                            unsigned long long* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
      ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(tm), "r"(x), "r"(y), "r"(z),
        "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <bool EDGE>
__device__ __forceinline__ void qg_tma_compute(const QgArgs<float>& A, const Stage<float>& st,
                                               const float (*s_q)[QHW], const float (*s_p)[QHW],
                                               const float (*s_epi)[QFY][QTXG * 4], int g0, int j0,
                                               int plane, int k) {
  const Layout& L = A.L;
  const int Ny = L.Ny, Nx = L.Nx;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int jA = j0 + QFR * ty, g = g0 + tx;
  const int r0 = QFR * ty, cs = 4 * (tx + 1);      // window rows r0 .. r0+3 of the halo tile
  float f[QFR + 2][6], q[QFR + 2][6];
#pragma unroll
  for (int d = 0; d < QFR + 2; ++d) {
    const Vec4<float> a = ld4(&s_p[r0 + d][cs]);
    const Vec4<float> b = ld4(&s_q[r0 + d][cs]);
    float al = __shfl_up_sync(0xffffffffu, a.w, 1), ar = __shfl_down_sync(0xffffffffu, a.x, 1);
    float bl = __shfl_up_sync(0xffffffffu, b.w, 1), br = __shfl_down_sync(0xffffffffu, b.x, 1);
    if (tx == 0) { al = s_p[r0 + d][cs - 1]; bl = s_q[r0 + d][cs - 1]; }
    if (tx == QTXG - 1) { ar = s_p[r0 + d][cs + 4]; br = s_q[r0 + d][cs + 4]; }
    f[d][0] = al; f[d][1] = a.x; f[d][2] = a.y; f[d][3] = a.z; f[d][4] = a.w; f[d][5] = ar;
    q[d][0] = bl; q[d][1] = b.x; q[d][2] = b.y; q[d][3] = b.z; q[d][4] = b.w; q[d][5] = br;
  }
  if (EDGE && (g >= L.groups() || jA >= Ny)) return;
  float bet[QFR + 2];
#pragma unroll
  for (int d = 0; d < QFR + 2; ++d) {
    int jj = jA - 1 + d;
    if (EDGE) jj = jj < 0 ? 0 : (jj > Ny - 1 ? Ny - 1 : jj);
    bet[d] = A.beta[jj];
  }
  const size_t po = (size_t)plane * L.plane();
  const int i0 = g * 4 - OFF;
#pragma unroll
  for (int rr = 0; rr < QFR; ++rr) {
    const int j = jA + rr;
    if (EDGE && j >= Ny) break;
    float wind = 0.f;
    if (k == 0) wind = (A.tau0 * A.wind[j]) * A.iH0;
    float out[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = i0 + e, c = e + 1;
      const int d0 = rr, d1 = rr + 1, d2 = rr + 2;      // window rows j-1, j, j+1
      float dq = 0.f;
      const bool inside = !EDGE || (i >= 0 && i < Nx);
      const bool interior = !EDGE || (j >= 1 && j <= Ny - 2 && i >= 1 && i <= Nx - 2);
      if (inside) {
        if (interior) {
          const float fE = f[d1][c + 1], fW = f[d1][c - 1], fN = f[d2][c], fS = f[d0][c];
          const float fNE = f[d2][c + 1], fNW = f[d2][c - 1], fSE = f[d0][c + 1], fSW = f[d0][c - 1];
          const float gE = q[d1][c + 1] + bet[d1], gW = q[d1][c - 1] + bet[d1];
          const float gN = q[d2][c] + bet[d2], gS = q[d0][c] + bet[d0];
          const float gNE = q[d2][c + 1] + bet[d2], gNW = q[d2][c - 1] + bet[d2];
          const float gSE = q[d0][c + 1] + bet[d0], gSW = q[d0][c - 1] + bet[d0];
          const float jpp = (fE - fW) * (gN - gS) - (fN - fS) * (gE - gW);
          const float jpx = fE * (gNE - gSE) - fW * (gNW - gSW) - fN * (gNE - gNW) + fS * (gSE - gSW);
          const float jxp = gN * (fNE - fNW) - gS * (fSE - fSW) - gE * (fNE - fSE) + gW * (fNW - fSW);
          dq = -(((jpp + jpx) + jxp) * A.ijden);
        }
        if (k == 0) dq = dq + wind;
        if (interior) {
          if (k == L.nl - 1) {
            const float pc = f[d1][c];
            const float lap = (f[d1][c + 1] - 2.f * pc + f[d1][c - 1]) * A.idx2 +
                              (f[d2][c] - 2.f * pc + f[d0][c]) * A.idy2;
            dq = dq + (-A.kappa * lap);
          }
          const float qc = q[d1][c];
          const float lapq = (q[d1][c + 1] - 2.f * qc + q[d1][c - 1]) * A.idx2 +
                             (q[d2][c] - 2.f * qc + q[d0][c]) * A.idy2;
          dq = dq + A.nu * lapq;
        }
      }
      out[e] = dq;
    }
    const size_t idx = po + (size_t)j * L.pitch + (size_t)g * 4;
    const Vec4<float> F{out[0], out[1], out[2], out[3]};
    if (st.Fout[0]) st4(st.Fout[0] + idx, F);
    if (st.Yout[0]) {
      const int er = QFR * ty + rr;
      Vec4<float> acc = ld4(&s_epi[MAX_PREV][er][4 * tx]);
#pragma unroll
      for (int jj = 0; jj < MAX_PREV; ++jj) {
        if (jj < st.nprev) {
          const Vec4<float> kk = ld4(&s_epi[jj][er][4 * tx]);
          acc.x = fmaf(st.adt[jj], kk.x, acc.x); acc.y = fmaf(st.adt[jj], kk.y, acc.y);
          acc.z = fmaf(st.adt[jj], kk.z, acc.z); acc.w = fmaf(st.adt[jj], kk.w, acc.w);
        }
      }
      acc.x = fmaf(st.adt_new, F.x, acc.x); acc.y = fmaf(st.adt_new, F.y, acc.y);
      acc.z = fmaf(st.adt_new, F.z, acc.z); acc.w = fmaf(st.adt_new, F.w, acc.w);
      st4(st.Yout[0] + idx, acc);
    }
  }
}

__global__ void __launch_bounds__(QTXG* QTY, 3)
qg_rhs_kernel_tma(const __grid_constant__ QgTmaps M, const QgArgs<float> A, const Stage<float> st) {
  extern __shared__ __align__(128) unsigned char qsm[];
  float (*s_q)[QHW] = reinterpret_cast<float (*)[QHW]>(qsm);
  float (*s_p)[QHW] = reinterpret_cast<float (*)[QHW]>(qsm + QG_TMA_HALO_B);
  float (*s_epi)[QFY][QTXG * 4] = reinterpret_cast<float (*)[QFY][QTXG * 4]>(qsm + 2 * QG_TMA_HALO_B);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(qsm + 2 * QG_TMA_HALO_B + (MAX_PREV + 1) * QG_TMA_EPI_B);
  const Layout& L = A.L;
  const int tid = threadIdx.y * QTXG + threadIdx.x;
  // Tiles start at the first interior cell (row 1, column 1 = group 1), so a grid whose interior is a
  // multiple of the tile (every power-of-two grid: 256^2 members, 8192^2) splits into whole,
  // predicate-free tiles; row 0 and column 0 - and the last row / column when no tile reaches them -
  // are written by qg_ring_kernel.  (With the origin at row 0 / group 0 a 256^2 member took 3 x 17
  // tiles, every one of them on the predicated path, instead of 2 x 16.)
  const int g0 = 1 + blockIdx.x * QTXG, j0 = 1 + blockIdx.y * QFY;
  const int plane = blockIdx.z;
  const int k = L.nl == 1 ? 0 : plane - (int)__umulhi((unsigned)plane, A.nl_magic) * L.nl;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    const bool epi = st.Yout[0] != nullptr;
    const unsigned bytes = 2u * QHR * QHW * 4u + (epi ? (unsigned)(1 + st.nprev) * (unsigned)QG_TMA_EPI_B : 0u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                 ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
    tma_load_3d(&s_q[0][0], &M.q, 4 * g0 - 4, j0 - 1, plane, bar);
    tma_load_3d(&s_p[0][0], &M.psi, 4 * g0 - 4, j0 - 1, plane, bar);
    if (epi) {
      tma_load_3d(&s_epi[MAX_PREV][0][0], &M.base, 4 * g0, j0, plane, bar);
#pragma unroll
      for (int jj = 0; jj < MAX_PREV; ++jj)
        if (jj < st.nprev) tma_load_3d(&s_epi[jj][0][0], &M.f[jj], 4 * g0, j0, plane, bar);
    }
  }
  __syncthreads();                      // barrier initialised before anyone waits on it
  {
    unsigned ok = 0;
    const unsigned ba = (unsigned)__cvta_generic_to_shared(bar);
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\t"
                   "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(ba) : "memory");
    }
  }
  // boundary condition on load: ring cells of q inside the halo tile are zero
  if (A.apply_bc) {
    const int rtop = L.Ny - 1 - (j0 - 1);                         // tile row of domain row Ny-1
    const int cright = L.Nx - 1 + OFF - 4 * (g0 - 1);             // tile column of domain column Nx-1
    const bool r0 = blockIdx.y == 0 && A.bc_ylo, r1 = rtop < QHR && A.bc_yhi;
    const bool c0 = blockIdx.x == 0, c1 = cright < QHW;
    if (r0 | r1 | c0 | c1) {
      if (r0 && tid < QHW) s_q[0][tid] = 0.f;
      if (r1 && tid < QHW) s_q[rtop][tid] = 0.f;
      if (c0 && tid < QHR) s_q[tid][OFF] = 0.f;
      if (c1 && tid < QHR) s_q[tid][cright] = 0.f;
      __syncthreads();
    }
  }
  // tiles that reach past the last interior row / column take the predicated path
  const bool edge = 4 * g0 + 4 * QTXG - 4 > L.Nx - 2 || j0 + QFY - 1 > L.Ny - 2;
  if (edge) qg_tma_compute<true>(A, st, s_q, s_p, s_epi, g0, j0, plane, k);
  else qg_tma_compute<false>(A, st, s_q, s_p, s_epi, g0, j0, plane, k);
}

// Ring cells no tile of qg_rhs_kernel_tma covers: row 0 and column 0 always, row Ny-1 / column Nx-1
// when the interior is a whole number of tiles in that direction.  Nothing but the wind acts there
// (same arithmetic as the predicated path of the tile kernel).  blockIdx.y: 0 = row 0, 1 = column 0,
// 2 = row Ny-1, 3 = column Nx-1; rows take the corners.
__global__ void qg_ring_kernel(const QgArgs<float> A, const Stage<float> st, int do_top, int do_right) {
  const Layout& L = A.L;
  const int t = blockIdx.x * blockDim.x + threadIdx.x, which = blockIdx.y, plane = blockIdx.z;
  int j, i;
  if (which == 0 || which == 2) {
    if (which == 2 && !do_top) return;
    j = which == 0 ? 0 : L.Ny - 1; i = t;
    if (i >= L.Nx) return;
  } else {
    if (which == 3 && !do_right) return;
    i = which == 1 ? 0 : L.Nx - 1; j = 1 + t;
    if (j > (do_top ? L.Ny - 2 : L.Ny - 1)) return;
  }
  const int k = L.nl == 1 ? 0 : plane % L.nl;
  float F = 0.f;
  if (k == 0) F = F + (A.tau0 * A.wind[j]) * A.iH0;
  const size_t idx = (size_t)plane * L.plane() + (size_t)j * L.pitch + OFF + i;
  if (st.Fout[0]) st.Fout[0][idx] = F;
  if (st.Yout[0]) {
    float acc = (st.y[0] ? st.y[0] : st.Yin[0])[idx];
#pragma unroll
    for (int jj = 0; jj < MAX_PREV; ++jj)
      if (jj < st.nprev) acc = fmaf(st.adt[jj], st.Fprev[jj][0][idx], acc);
    acc = fmaf(st.adt_new, F, acc);
    st.Yout[0][idx] = acc;
  }
}

// ring := 0 in place on padded planes
template <typename T>
__global__ void qg_bc_kernel(T* __restrict__ q, Layout L, int bc_ylo, int bc_yhi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y, p = blockIdx.z;
  if (i >= L.Nx) return;
  if ((j == 0 && bc_ylo) || (j == L.Ny - 1 && bc_yhi) || i == 0 || i == L.Nx - 1)
    q[(size_t)p * L.plane() + (size_t)j * L.pitch + OFF + i] = T(0);
}

// KE / enstrophy / non-finite count from q (reference layout) and psi (padded layout).
template <typename T>
__global__ void qg_diag_kernel(const T* __restrict__ q, const T* __restrict__ psi, Layout L,
                               double dx, double dy, double* __restrict__ out) {
  const int plane = blockIdx.z, b = plane / L.nl, k = plane % L.nl;
  const T* Q = q + (size_t)plane * L.Ny * L.Nx;
  const T* P = psi + (size_t)plane * L.plane();
  const int Ny = L.Ny, Nx = L.Nx, pitch = L.pitch;
  double ke = 0, ens = 0, bad = 0;
  for (int j = blockIdx.y; j < Ny; j += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nx; i += gridDim.x * blockDim.x) {
      T qv = Q[(size_t)j * Nx + i];
      if (!isfinite((double)qv)) bad += 1;
      if (j >= 1 && j <= Ny - 2 && i >= 1 && i <= Nx - 2) {
        auto PS = [&](int jj, int ii) { return P[(size_t)jj * pitch + OFF + ii]; };
        // u = -d psi/dy at V points (interior-only, zero ring), v = d psi/dx at U points
        auto uV = [&](int jj, int ii) -> T {
          return (jj >= 1 && jj <= Ny - 2 && ii >= 1 && ii <= Nx - 2)
                     ? -((PS(jj + 1, ii) - PS(jj, ii)) / (T)dy) : T(0); };
        auto vU = [&](int jj, int ii) -> T {
          return (jj >= 1 && jj <= Ny - 2 && ii >= 1 && ii <= Nx - 2)
                     ? (PS(jj, ii + 1) - PS(jj, ii)) / (T)dx : T(0); };
        T uT = T(0.5) * (uV(j, i) + uV(j - 1, i));
        T vT = T(0.5) * (vU(j, i) + vU(j, i - 1));
        ke += (double)(uT * uT + vT * vT);
        ens += (double)(qv * qv);
      }
    }
  __shared__ double red[3][8];
  double vals[3] = {ke, ens, bad};
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int qd = 0; qd < 3; ++qd) {
    double x = vals[qd];
    for (int s = 16; s > 0; s >>= 1) x += __shfl_down_sync(0xffffffffu, x, s);
    if (lane == 0) red[qd][w] = x;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double x = 0;
    for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) x += red[threadIdx.x][ww];
    double area = dx * dy;
    double* o = out + (size_t)b * (2 * L.nl + 1);
    if (threadIdx.x == 0) atomicAdd(o + k, 0.5 * x * area);
    if (threadIdx.x == 1) atomicAdd(o + L.nl + k, 0.5 * x * area);
    if (threadIdx.x == 2) atomicAdd(o + 2 * L.nl, x);
  }
}

}  // namespace sb

using namespace sb;

struct somax_b200_qg_s {
  int dtype;
  Layout L;
  int ny, nx;
  double dx, dy;
  unsigned spec;
  QgSolver* solver = nullptr;
  int bc_ylo = 1, bc_yhi = 1;   // physical ring rows (a slab of the distributed model clears them)
  void* beta = nullptr; void* wind = nullptr;
  bool beta1d = false, wind1d = false;
  void* y = nullptr; void* Ya = nullptr; void* Yb = nullptr; void* psi = nullptr;
  void* F[5] = {0, 0, 0, 0, 0};
  // tensor maps of the 9 internal arrays for the TMA stencil (fp32): halo box and epilogue box
  struct TmEntry { const void* ptr; CUtensorMap halo, epi; };
  TmEntry tm[9];
  int ntm = 0;
  StepGraph graph;
  size_t bytes = 0;
};

namespace {

// BarotropicQG._invert_pv returns the solver's output as it is (qg/barotropic.py:113-121);
// BaroclinicQG._invert_pv zeroes the ring of psi (qg/baroclinic.py:157-158).
inline int keep_psi_ring(const somax_b200_qg_s* h) { return (h->spec & SOMAX_B200_SPEC_KEEP_PSI_RING) ? 1 : 0; }

template <typename T>
QgArgs<T> make_qargs(somax_b200_qg_t h, const somax_b200_params* p, int apply_bc) {
  QgArgs<T> A;
  A.L = h->L; A.apply_bc = apply_bc; A.bc_ylo = h->bc_ylo; A.bc_yhi = h->bc_yhi;
  A.dx2 = (T)(h->dx * h->dx); A.dy2 = (T)(h->dy * h->dy); A.jden = (T)(12.0 * h->dx * h->dy);
  A.idx2 = (T)(1.0 / (h->dx * h->dx)); A.idy2 = (T)(1.0 / (h->dy * h->dy));
  A.ijden = (T)(1.0 / (12.0 * h->dx * h->dy)); A.iH0 = (T)(1.0 / p->H0);
  A.beta = (const T*)h->beta; A.b_cp = h->beta1d ? 1 : h->L.Nx; A.b_xs = h->beta1d ? 0 : 1;
  A.wind = (const T*)h->wind; A.w_cp = h->wind1d ? 1 : h->L.Nx; A.w_xs = h->wind1d ? 0 : 1;
  A.H0 = (T)p->H0; A.nu = (T)p->lateral_viscosity; A.kappa = (T)p->bottom_drag;
  A.tau0 = (T)p->wind_amplitude;
  A.nl_magic = h->L.nl > 1 ? (unsigned)(((1ull << 32) + h->L.nl - 1) / h->L.nl) : 0u;
  return A;
}

template <typename T>
Stage<T> qstage() {
  Stage<T> st;
  st.nfields = 1; st.nprev = 0;
  for (int f = 0; f < MAX_FIELDS; ++f) {
    st.Yin[f] = nullptr; st.y[f] = nullptr; st.Fout[f] = nullptr; st.Yout[f] = nullptr;
    for (int j = 0; j < MAX_PREV; ++j) st.Fprev[j][f] = nullptr;
  }
  for (int j = 0; j < MAX_PREV; ++j) st.a[j] = 0;
  st.a_new = 0; st.dt = 0;
  return st;
}

// Tensor maps over a padded (plane, Ny, pitch) fp32 array; OOB elements read as zero.
int qg_build_tmaps(somax_b200_qg_t h) {
  h->ntm = 0;
  if (h->dtype != SOMAX_B200_F32 || !(h->beta1d && h->wind1d) || getenv("SOMAX_B200_NO_TMA")) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !fn) {
    cudaGetLastError();
    return 0;                       // fall back to the cp.async kernel
  }
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
  const Layout& L = h->L;
  const cuuint64_t dims[3] = {(cuuint64_t)L.pitch, (cuuint64_t)L.Ny, (cuuint64_t)L.batch * L.nl};
  const cuuint64_t strides[2] = {(cuuint64_t)L.pitch * 4, (cuuint64_t)L.plane() * 4};
  const cuuint32_t estr[3] = {1, 1, 1};
  const cuuint32_t box_halo[3] = {(cuuint32_t)QHW, (cuuint32_t)QHR, 1};
  const cuuint32_t box_epi[3] = {(cuuint32_t)(QTXG * 4), (cuuint32_t)QFY, 1};
  void* bufs[9] = {h->y, h->Ya, h->Yb, h->psi, h->F[0], h->F[1], h->F[2], h->F[3], h->F[4]};
  for (int i = 0; i < 9; ++i) {
    somax_b200_qg_s::TmEntry& e = h->tm[i];
    e.ptr = bufs[i];
    CUresult r1 = encode(&e.halo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, bufs[i], dims, strides, box_halo, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = encode(&e.epi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, bufs[i], dims, strides, box_epi, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { h->ntm = 0; return 0; }
  }
  h->ntm = 9;
  return 0;
}

const somax_b200_qg_s::TmEntry* qg_find_tm(somax_b200_qg_t h, const void* p) {
  for (int i = 0; i < h->ntm; ++i)
    if (h->tm[i].ptr == p) return &h->tm[i];
  return nullptr;
}

// fp32 fast path through the TMA kernel; returns false when some operand is not an internal array
bool qg_launch_tma(somax_b200_qg_t h, const QgArgs<float>& A, const Stage<float>& st, cudaStream_t s) {
  if (h->ntm == 0) return false;
  QgTmaps M;
  const auto* eq = qg_find_tm(h, st.Yin[0]);
  const auto* ep = qg_find_tm(h, h->psi);
  if (!eq || !ep) return false;
  M.q = eq->halo; M.psi = ep->halo;
  M.base = eq->epi;
  for (int jj = 0; jj < MAX_PREV; ++jj) M.f[jj] = eq->epi;
  if (st.Yout[0]) {
    const auto* eb = qg_find_tm(h, st.y[0] ? st.y[0] : st.Yin[0]);
    if (!eb) return false;
    M.base = eb->epi;
    for (int jj = 0; jj < st.nprev; ++jj) {
      const auto* ef = qg_find_tm(h, st.Fprev[jj][0]);
      if (!ef) return false;
      M.f[jj] = ef->epi;
    }
  }
  if (ensure_dyn_smem((const void*)qg_rhs_kernel_tma, QG_TMA_SMEM) != 0) { cudaGetLastError(); return false; }
  const Layout& L = h->L;
  dim3 block(QTXG, QTY);
  // tiles over the interior rows 1 .. Ny-2 and the groups 1 .. gI of the interior columns
  const int gI = (L.Nx - 2 + OFF) >> 2, gR = (L.Nx - 1 + OFF) >> 2;
  const int gx = (gI + QTXG - 1) / QTXG, gy = (L.Ny - 2 + QFY - 1) / QFY;
  dim3 grid(gx, gy, L.batch * L.nl);
  prof_begin("qg_rhs_kernel", s);
  qg_rhs_kernel_tma<<<grid, block, QG_TMA_SMEM, s>>>(M, A, st);
  const int do_top = QFY * gy < L.Ny - 1, do_right = gR > QTXG * gx;
  const int nmax = L.Nx > L.Ny ? L.Nx : L.Ny;
  prof_end();
  g_launches.fetch_add(1, std::memory_order_relaxed);
  prof_begin("qg_ring_kernel", s);
  qg_ring_kernel<<<dim3((nmax + 255) / 256, 4, L.batch * L.nl), 256, 0, s>>>(A, st, do_top, do_right);
  return true;
}

template <typename T>
int launch_stencil(somax_b200_qg_t h, const QgArgs<T>& A, const Stage<T>& st_in, double dt, cudaStream_t s);

// one RHS evaluation: psi = invert(Yin), then the fused stencil + RK epilogue
template <typename T>
int eval_rhs(somax_b200_qg_t h, const QgArgs<T>& A, const Stage<T>& st_in, double dt, cudaStream_t s) {
  if (int rc = qg_solver_run<T>(h->solver, st_in.Yin[0], (T*)h->psi, A.apply_bc, keep_psi_ring(h), s)) return rc;
  return launch_stencil<T>(h, A, st_in, dt, s);
}

// the fused stencil + RK epilogue alone (psi already in h->psi)
template <typename T>
int launch_stencil(somax_b200_qg_t h, const QgArgs<T>& A, const Stage<T>& st_in, double dt, cudaStream_t s) {
  Stage<T> st = st_in;
  stage_finalize(st, dt);
  const Layout& L = h->L;
  dim3 block(QTXG, QTY);
  dim3 grid((L.groups() + QTXG - 1) / QTXG, (L.Ny + QTY - 1) / QTY, L.batch * L.nl);
  if constexpr (sizeof(T) == 4) {
    if (qg_launch_tma(h, A, st, s)) { SB_LAUNCH_CHECK(); return 0; }
  }
  prof_begin("qg_rhs_kernel", s);
  if (h->beta1d && h->wind1d) qg_rhs_kernel_fast<T><<<grid, block, 0, s>>>(A, (const T*)h->psi, st);
  else qg_rhs_kernel<T><<<grid, block, 0, s>>>(A, (const T*)h->psi, st);
  SB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int qg_bc_inplace(somax_b200_qg_t h, void* q, cudaStream_t s) {
  const Layout& L = h->L;
  dim3 b(256), g((L.Nx + 255) / 256, L.Ny, L.batch * L.nl);
  prof_begin("qg_bc_kernel", s);
  qg_bc_kernel<T><<<g, b, 0, s>>>((T*)q, L, h->bc_ylo, h->bc_yhi);
  SB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
int qg_steps_impl(somax_b200_qg_t h, void* q, long n_steps, double dt, double dt_last,
                  const somax_b200_params* p, bool bc0, cudaStream_t caller) {
  const Layout& L = h->L;
  const long total = n_steps + (dt_last > 0 ? 1 : 0);
  // launch-bound grids: replay two-step CUDA graphs on an internal stream
  const bool use_graph = !prof_enabled() && L.count() <= GRAPH_MAX_CELLS && n_steps >= 9 &&
                         h->graph.init() == 0;
  cudaStream_t s = caller;
  if (use_graph) {
    s = h->graph.stream;
    SB_CUDA(cudaEventRecord(h->graph.ev_in, caller));
    SB_CUDA(cudaStreamWaitEvent(s, h->graph.ev_in, 0));
  }
  void *y = h->y, *Yc = h->Ya, *Yn = h->Yb;
  if (int rc = pack_field<T>((const T*)q, (T*)y, L, s)) return rc;
  if (bc0)
    if (int rc = qg_bc_inplace<T>(h, y, s)) return rc;   // integrate(): BC on state0 (not when resuming)
  if (total > 0) {
    QgArgs<T> A = make_qargs<T>(h, p, 1);
    auto step_dt = [&](long i) { return (i < n_steps) ? dt : dt_last; };
    {
      Stage<T> st = qstage<T>();
      st.Yin[0] = (const T*)y; st.Fout[0] = (T*)h->F[0]; st.Yout[0] = (T*)Yc;
      st.a_new = (T)TSIT5_A[0][0]; st.dt = (T)step_dt(0);
      if (int rc = eval_rhs<T>(h, A, st, step_dt(0), s)) return rc;
    }
    // stages 2..6 of a step (evaluations at Y2..Y6, forming Y7 = y_{n+1} in Yc)
    auto stages = [&](double hd) -> int {
      for (int e = 1; e <= 5; ++e) {
        Stage<T> st = qstage<T>();
        st.nprev = e; st.dt = (T)hd; st.a_new = (T)TSIT5_A[e][e];
        for (int jj = 0; jj < e; ++jj) { st.a[jj] = (T)TSIT5_A[e][jj]; st.Fprev[jj][0] = (const T*)h->F[jj]; }
        st.Yin[0] = (const T*)Yc; st.y[0] = (const T*)y; st.Yout[0] = (T*)Yn;
        st.Fout[0] = (e <= 4) ? (T*)h->F[e] : nullptr;
        if (int rc = eval_rhs<T>(h, A, st, hd, s)) return rc;
        std::swap(Yc, Yn);
      }
      return 0;
    };
    // a step followed by another one: also the FSAL evaluation f(Y7) = next step's F1 and Y2
    auto full_step = [&](double hd, double hnext) -> int {
      if (int rc = stages(hd)) return rc;
      Stage<T> st = qstage<T>();
      st.dt = (T)hnext; st.a_new = (T)TSIT5_A[0][0];
      st.Yin[0] = (const T*)Yc; st.Fout[0] = (T*)h->F[0]; st.Yout[0] = (T*)Yn;
      if (int rc = eval_rhs<T>(h, A, st, hnext, s)) return rc;
      void* oy = y; y = Yc; Yc = Yn; Yn = oy;
      return 0;
    };
    long i = 0;
    if (use_graph) {
      StepGraph& G = h->graph;
      const double key[6] = {dt, p->lateral_viscosity, p->bottom_drag, p->wind_amplitude, p->H0, 1.0};
      bool same = G.exec != nullptr;
      for (int k = 0; k < 6; ++k) same = same && (G.key[k] == key[k]);
      if (!same) {
        if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
        const uint64_t n0 = g_launches.load();
        cudaGraph_t graph = nullptr;
        SB_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
        int rc = full_step(dt, dt);
        if (!rc) rc = full_step(dt, dt);          // buffers are back in their starting roles
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail(SOMAX_B200_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&G.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) return fail(SOMAX_B200_ERR_CUDA, std::string("graph instantiate: ") + cudaGetErrorString(ce));
        G.nlaunch = g_launches.load() - n0;
        g_launches.fetch_sub(G.nlaunch);          // captured, not executed
        for (int k = 0; k < 6; ++k) G.key[k] = key[k];
      }
      const long pairs = (n_steps - 1) / 2;       // steps 0 .. n_steps-2 have dt before and after
      for (long r = 0; r < pairs; ++r) SB_CUDA(cudaGraphLaunch(G.exec, s));
      g_launches.fetch_add(G.nlaunch * (uint64_t)pairs);
      i = 2 * pairs;
    }
    for (; i + 1 < total; ++i)
      if (int rc = full_step(step_dt(i), step_dt(i + 1))) return rc;
    if (int rc = stages(step_dt(total - 1))) return rc;
    std::swap(y, Yc);
  }
  if (int rc = unpack_field<T>((const T*)y, (T*)q, L, s)) return rc;
  if (use_graph) {
    SB_CUDA(cudaEventRecord(h->graph.ev_out, s));
    SB_CUDA(cudaStreamWaitEvent(caller, h->graph.ev_out, 0));
  }
  return 0;
}

template <typename T>
int qg_rhs_impl(somax_b200_qg_t h, const void* q, void* dq, void* psi_out,
                const somax_b200_params* p, int apply_bc, cudaStream_t s) {
  const Layout& L = h->L;
  if (int rc = pack_field<T>((const T*)q, (T*)h->Ya, L, s)) return rc;
  QgArgs<T> A = make_qargs<T>(h, p, apply_bc);
  Stage<T> st = qstage<T>();
  st.Yin[0] = (const T*)h->Ya; st.Fout[0] = (T*)h->F[0];
  if (int rc = eval_rhs<T>(h, A, st, 0.0, s)) return rc;
  if (int rc = unpack_field<T>((const T*)h->F[0], (T*)dq, L, s)) return rc;
  if (psi_out) return unpack_field<T>((const T*)h->psi, (T*)psi_out, L, s);
  return 0;
}

template <typename T>
int qg_invert_impl(somax_b200_qg_t h, const void* q, void* psi, cudaStream_t s) {
  const Layout& L = h->L;
  if (int rc = pack_field<T>((const T*)q, (T*)h->Ya, L, s)) return rc;
  if (int rc = qg_solver_run<T>(h->solver, (const T*)h->Ya, (T*)h->psi, 0, keep_psi_ring(h), s)) return rc;
  return unpack_field<T>((const T*)h->psi, (T*)psi, L, s);
}

template <typename T>
int qg_bc_impl(somax_b200_qg_t h, const void* q, void* out, cudaStream_t s) {
  const Layout& L = h->L;
  if (int rc = pack_field<T>((const T*)q, (T*)h->Ya, L, s)) return rc;
  if (int rc = qg_bc_inplace<T>(h, h->Ya, s)) return rc;
  return unpack_field<T>((const T*)h->Ya, (T*)out, L, s);
}

template <typename T>
int qg_diag_impl(somax_b200_qg_t h, const void* q, double* out, cudaStream_t s) {
  const Layout& L = h->L;
  if (int rc = pack_field<T>((const T*)q, (T*)h->Ya, L, s)) return rc;
  if (int rc = qg_solver_run<T>(h->solver, (const T*)h->Ya, (T*)h->psi, 0, keep_psi_ring(h), s)) return rc;
  SB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * L.batch * (2 * L.nl + 1), s));
  dim3 b(256), g(std::min((L.Nx + 255) / 256, 64), std::min(L.Ny, 128), L.batch * L.nl);
  prof_begin("qg_diag_kernel", s);
  qg_diag_kernel<T><<<g, b, 0, s>>>((const T*)q, (const T*)h->psi, L, h->dx, h->dy, out);
  SB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

#define SB_DISPATCH(h, fn, ...) \
  ((h)->dtype == SOMAX_B200_F32 ? fn<float>(__VA_ARGS__) : fn<double>(__VA_ARGS__))

// rows > 0: the PV-inversion row stage covers window rows [jo, jo + rows) only (slab of the
// distributed model, qg_slab.cuh); default: the whole array.
static int qg_create_impl(somax_b200_qg_t* out, int dtype, int batch, int nl, int ny, int nx,
                          double dx, double dy, const double* Cl2m, const double* Cm2l,
                          const double* lambdas, const double* beta_y, const double* wind,
                          int solver, unsigned spec_flags, int rows, int jo, int ylo, int yhi) {
  if (!out) return fail(SOMAX_B200_ERR_INVALID, "out is null");
  *out = nullptr;
  if (dtype != SOMAX_B200_F32 && dtype != SOMAX_B200_F64)
    return fail(SOMAX_B200_ERR_INVALID, "dtype must be F32 or F64");
  if (batch < 1 || nl < 1 || ny < 3 || nx < 3)
    return fail(SOMAX_B200_ERR_INVALID, "need batch>=1, nl>=1, ny>=3, nx>=3");
  if (!Cl2m || !Cm2l || !lambdas || !beta_y || !wind)
    return fail(SOMAX_B200_ERR_INVALID, "null coefficient pointer");
  if ((long)batch * nl > 65535) return fail(SOMAX_B200_ERR_UNSUPPORTED, "batch*nl > 65535");
  if (spec_flags & (SOMAX_B200_SPEC_DST_CONTINUOUS | SOMAX_B200_SPEC_DST_INTERIOR))
    return fail(SOMAX_B200_ERR_UNSUPPORTED,
                "only the reference's DST convention is implemented: finite-difference eigenvalues on the whole array");
  if (int rc = require_device()) return rc;
  auto* h = new somax_b200_qg_s();
  h->dtype = dtype; h->L = make_layout(batch, nl, ny, nx); h->ny = ny; h->nx = nx;
  h->dx = dx; h->dy = dy; h->spec = spec_flags;
  int rc = qg_solver_create(&h->solver, dtype, batch, nl, ny, nx, dx, dy, Cl2m, Cm2l, lambdas, solver, rows > 0 ? 1 : 0, rows, jo, ylo, yhi);
  auto up = [&](const double* src, void** dst, bool* one) {
    return dtype == SOMAX_B200_F32 ? upload_coef<float>(src, (float**)dst, h->L.Ny, h->L.Nx, one)
                                   : upload_coef<double>(src, (double**)dst, h->L.Ny, h->L.Nx, one);
  };
  if (!rc) rc = up(beta_y, &h->beta, &h->beta1d);
  if (!rc) rc = up(wind, &h->wind, &h->wind1d);
  const size_t fb = h->L.count() * (dtype == SOMAX_B200_F32 ? 4 : 8);
  void** bufs[] = {&h->y, &h->Ya, &h->Yb, &h->psi, &h->F[0], &h->F[1], &h->F[2], &h->F[3], &h->F[4]};
  for (void** bp : bufs) {
    if (rc) break;
    cudaError_t e = cudaMalloc(bp, fb);
    if (e != cudaSuccess) { rc = fail(SOMAX_B200_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); break; }
    cudaMemset(*bp, 0, fb);
    h->bytes += fb;
  }
  if (!rc) rc = qg_build_tmaps(h);
  if (rc) { somax_b200_qg_destroy(h); return rc; }
  h->bytes += qg_solver_bytes(h->solver);
  *out = h;
  return 0;
}

extern "C" {

int somax_b200_qg_create(somax_b200_qg_t* out, int dtype, int batch, int nl, int ny, int nx,
                         double dx, double dy, const double* Cl2m, const double* Cm2l,
                         const double* lambdas, const double* beta_y, const double* wind,
                         int solver, unsigned spec_flags) {
  return qg_create_impl(out, dtype, batch, nl, ny, nx, dx, dy, Cl2m, Cm2l, lambdas, beta_y, wind,
                        solver, spec_flags, -1, 0, 1, 1);
}

int somax_b200_qg_destroy(somax_b200_qg_t h) {
  if (!h) return 0;
  h->graph.destroy();
  qg_solver_destroy(h->solver);
  cudaFree(h->beta); cudaFree(h->wind);
  cudaFree(h->y); cudaFree(h->Ya); cudaFree(h->Yb); cudaFree(h->psi);
  for (int j = 0; j < 5; ++j) cudaFree(h->F[j]);
  delete h;
  return 0;
}

size_t somax_b200_qg_device_bytes(somax_b200_qg_t h) { return h ? h->bytes : 0; }

int somax_b200_qg_apply_bc(somax_b200_qg_t h, const void* q, void* out, void* stream) {
  if (!h || !q || !out) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  return SB_DISPATCH(h, qg_bc_impl, h, q, out, (cudaStream_t)stream);
}

int somax_b200_qg_invert(somax_b200_qg_t h, const void* q, void* psi, void* stream) {
  if (!h || !q || !psi) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  return SB_DISPATCH(h, qg_invert_impl, h, q, psi, (cudaStream_t)stream);
}

int somax_b200_qg_rhs(somax_b200_qg_t h, const void* q, void* dq, void* psi_out,
                      const somax_b200_params* p, int apply_bc, void* stream) {
  if (!h || !q || !dq || !p) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  return SB_DISPATCH(h, qg_rhs_impl, h, q, dq, psi_out, p, apply_bc, (cudaStream_t)stream);
}

int somax_b200_qg_steps(somax_b200_qg_t h, void* q, long n_steps, double dt, double dt_last,
                        const somax_b200_params* p, void* stream) {
  if (!h || !q || !p) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (n_steps < 0 || !(dt > 0) || dt_last < 0) return fail(SOMAX_B200_ERR_INVALID, "need n_steps>=0, dt>0, dt_last>=0");
  return SB_DISPATCH(h, qg_steps_impl, h, q, n_steps, dt, dt_last, p, true, (cudaStream_t)stream);
}

int somax_b200_qg_resume(somax_b200_qg_t h, void* q, long n_steps, double dt, double dt_last,
                         const somax_b200_params* p, void* stream) {
  if (!h || !q || !p) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (n_steps < 0 || !(dt > 0) || dt_last < 0) return fail(SOMAX_B200_ERR_INVALID, "need n_steps>=0, dt>0, dt_last>=0");
  return SB_DISPATCH(h, qg_steps_impl, h, q, n_steps, dt, dt_last, p, false, (cudaStream_t)stream);
}

int somax_b200_qg_diag(somax_b200_qg_t h, const void* q, double* out, void* stream) {
  if (!h || !q || !out) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  return SB_DISPATCH(h, qg_diag_impl, h, q, out, (cudaStream_t)stream);
}

}  // extern "C"

#include "qg_slab.cuh"
