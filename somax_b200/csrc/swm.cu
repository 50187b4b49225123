// Shallow-water hot path: single-pass fused RHS (mass flux, PV, Bernoulli, Coriolis
// interpolation, diffusion, wind, drag) with the Tsit5 stage combination in the epilogue.
//
// Replaces MultilayerShallowWater2D.{apply_boundary_conditions,vector_field} and the diffrax
// Tsit5 loop around them (reference swm/multilayer.py:150-223, core/model.py:47-88; operator
// semantics SURVEY.md App. B).  One kernel launch per RHS evaluation; every second-level
// operand (q, uh, vh, ke, P) is recomputed on chip from a shared-memory tile of (h,u,v) with a
// one-cell halo, reproducing finitevolx's "interior-only, zero ghost ring" intermediates by
// predicate on the global index.
#include "common.cuh"

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "swm_kernels.cuh"
#include "qg_solver.cuh"

namespace sb {

// ------------------------------------------------------------------------------------------
// Geostrophic projection of the reparameterized QG model, P = G . (Q.G)^-1 . Q
// (reference qg/reparameterized.py:142-189), applied by its apply_boundary_conditions after the
// shallow-water BCs.  Q and G are two small stencil kernels around the PV-inversion solver of the
// QG models (qg_solver.cu).
// ------------------------------------------------------------------------------------------
struct ProjArgs {
  double f0;
  double H[SWM_MAX_NL];
  double A[SWM_MAX_NL][SWM_MAX_NL];   // Cm2l . diag(eigenvalues) . Cl2m
};

// Q: q = curl(u, v) - f0 (h - H_k) / H_k on the whole array; the curl is finitevolx's
// relative_vorticity (interior only, zero ring), h - H_k is taken everywhere (reparameterized.py:160-164).
template <typename T>
__global__ void rqg_q_kernel(const T* __restrict__ h, const T* __restrict__ u, const T* __restrict__ v,
                             T* __restrict__ q, Layout L, int bc, int apply_bc, ProjArgs P, T dx, T dy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, plane = blockIdx.z;
  if (i >= L.Nx) return;
  const int k = plane % L.nl, Ny = L.Ny, Nx = L.Nx, pitch = L.pitch;
  const size_t po = (size_t)plane * L.plane();
  auto val = [&](const T* f, int kind, int jj, int ii) -> T {
    return apply_bc ? swm_bc_value(f + po, kind, bc, jj, ii, Ny, Nx, pitch) : f[po + (size_t)jj * pitch + OFF + ii];
  };
  T zeta = 0;
  if (j >= 1 && j <= Ny - 2 && i >= 1 && i <= Nx - 2)
    zeta = (val(v, FV, j, i + 1) - val(v, FV, j, i)) / dx - (val(u, FU, j + 1, i) - val(u, FU, j, i)) / dy;
  const T Hk = (T)P.H[k];
  const T eta = val(h, FH, j, i) - Hk;
  q[po + (size_t)j * pitch + OFF + i] = zeta - ((T)P.f0 * eta) / Hk;
}

// G: (u_g, v_g) = grad_perp(psi) = (-d psi/dy at U points, d psi/dx at V points), interior only with
// a zero ring; h_g = H_k (1 + f0 (A psi)_k), A psi = Cm2l (eigenvalues * (Cl2m psi)), everywhere
// (reparameterized.py:169-177).
template <typename T>
__global__ void rqg_g_kernel(const T* __restrict__ psi, T* __restrict__ hg, T* __restrict__ ug,
                             T* __restrict__ vg, Layout L, ProjArgs P, T dx, T dy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, plane = blockIdx.z;
  if (i >= L.Nx) return;
  const int b = plane / L.nl, k = plane - b * L.nl, Ny = L.Ny, Nx = L.Nx, pitch = L.pitch;
  const size_t o = (size_t)j * pitch + OFF + i;
  const size_t po = (size_t)plane * L.plane();
  T uu = 0, vv = 0;
  if (j >= 1 && j <= Ny - 2 && i >= 1 && i <= Nx - 2) {
    const T pc = psi[po + o];
    uu = -((pc - psi[po + o - pitch]) / dy);
    vv = (pc - psi[po + o - 1]) / dx;
  }
  T ap = 0;
  for (int c = 0; c < L.nl; ++c) ap += (T)P.A[k][c] * psi[((size_t)b * L.nl + c) * L.plane() + o];
  const T Hk = (T)P.H[k];
  hg[po + o] = Hk * (T(1) + (T)P.f0 * ap);
  ug[po + o] = uu;
  vg[po + o] = vv;
}

// Fast variant (f, wind_x, wind_y depend on y only - every reference factory): 128-bit tile
// loads, per-thread 3x6 register windows filled with 128-bit shared reads + warp shuffles,
// shared sub-expressions (q, uh, vh, ke, fluxes) evaluated once per window column, reciprocal
// multiplies for the constant divisors.
constexpr int SSW = TXG * 4 + 8;   // shared row: 4 | 128 | 4 (halo columns at [3] and [132])

// EDGE = false: the CTA's tile and its halo lie inside the region where every load is a plain load,
// every window cell is interior and every output is written - 97 % of the CTAs of a 4096^2 grid; that
// instance carries no ring / range predicates at all (they were a quarter of the instructions of a
// kernel that is issue bound, not HBM bound).
__device__ __forceinline__ float swm_div(float a, float b) { return __fdividef(a, b); }     // <= 2 ulp
__device__ __forceinline__ double swm_div(double a, double b) { return a / b; }

template <typename T, bool EDGE>
__device__ __forceinline__ void swm_rhs_body(const SwmArgs<T>& A, const Stage<T>& st, int j0, T (*s_h)[SSW],
                                             T (*s_u)[SSW], T (*s_v)[SSW], T (*s_p)[SSW]) {
  const Layout& L = A.L;
  const int Ny = L.Ny, Nx = L.Nx, pitch = L.pitch, ngroups = L.groups();
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TXG + tx;
  const int g0 = blockIdx.x * TXG;
  const int b = blockIdx.z;
  const int j = j0 + ty, g = g0 + tx;
  const int r = ty + 1, cs = 4 * (tx + 1);
  const int i0 = g * 4 - OFF;                    // column of this thread's first cell
  const T idx_ = A.idx, idy_ = A.idy;
  const T nux = A.nu * idx_, nuy = A.nu * idy_;      // as in swm_rhs_inner: the two paths agree bit for bit

  for (int e = tid; e < (TY + 2) * (SSW / 4); e += TXG * TY)
    st4(&s_p[0][0] + 4 * e, Vec4<T>{0, 0, 0, 0});

  // row / column interior flags of the 3 x 6 window
  bool rin[3], cin[6];
#pragma unroll
  for (int d = 0; d < 3; ++d) rin[d] = !EDGE || swm_row_interior(j - 1 + d, Ny, A.ylo, A.yhi);
#pragma unroll
  for (int w = 0; w < 6; ++w) cin[w] = !EDGE || ((i0 - 1 + w >= 1) && (i0 - 1 + w <= Nx - 2));
  // Coriolis at the X points of rows j-1 and j: f depends on y only, so
  // T_to_X(f) = 0.25*(f[j] + f[j] + f[j+1] + f[j+1]) in the reference's summation order
  T fX[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    int ja = j - 1 + d, jb = ja + 1;
    if (EDGE) { ja = ja < 0 ? 0 : (ja > Ny - 1 ? Ny - 1 : ja); jb = ja + 1 > Ny - 1 ? Ny - 1 : ja + 1; }
    const T fa = A.f[ja], fb = A.f[jb];
    fX[d] = T(0.25) * (((fa + fa) + fb) + fb);
  }
  const int jc = (!EDGE || j < Ny) ? j : Ny - 1;
  const T windx = (A.tau0 * A.wx[jc]) * A.iH0, windy = (A.tau0 * A.wy[jc]) * A.iH0;

  // fp32: the Tsit5 epilogue operands of this layer (step-start state and previous stage
  // derivatives of h, u, v: up to 18 vectors per thread) are prefetched with 16-byte cp.async into
  // per-thread shared slots before the tile load, so their HBM latency hides behind the tile
  // load, the barriers and the stencil arithmetic instead of sitting in front of the final FMAs.
  constexpr bool STAGE_EPI = sizeof(T) == 4;
  extern __shared__ __align__(16) unsigned char swm_dyn[];
  T (*s_epi)[TXG * TY][4] = reinterpret_cast<T (*)[TXG * TY][4]>(swm_dyn);     // [3 * (MAX_PREV + 1)]
  const bool epi_valid = (!EDGE || (j < Ny && g < ngroups)) && st.Yout[FH] != nullptr;

  for (int k = 0; k < L.nl; ++k) {
    const size_t plane_off = ((size_t)b * L.nl + k) * L.plane();
    const T* ph = st.Yin[FH] + plane_off;
    const T* pu = st.Yin[FU] + plane_off;
    const T* pv = st.Yin[FV] + plane_off;
    // periodic basin in slabs: source rows of the physical ghost rows (SwmArgs::lo_src / hi_src)
    const size_t ro = ((size_t)b * L.nl + k) * pitch;
    const T* loh = A.lo_src[FH] ? A.lo_src[FH] + ro : nullptr; const T* hih = A.hi_src[FH] ? A.hi_src[FH] + ro : nullptr;
    const T* lou = A.lo_src[FU] ? A.lo_src[FU] + ro : nullptr; const T* hiu = A.hi_src[FU] ? A.hi_src[FU] + ro : nullptr;
    const T* lov = A.lo_src[FV] ? A.lo_src[FV] + ro : nullptr; const T* hiv = A.hi_src[FV] ? A.hi_src[FV] + ro : nullptr;
    if (STAGE_EPI) {
      if (epi_valid) {
        const size_t eidx = plane_off + (size_t)j * pitch + (size_t)g * 4;
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          const T* base = (st.y[f] ? st.y[f] : st.Yin[f]) + eidx;
          unsigned sa = (unsigned)__cvta_generic_to_shared(&s_epi[f * (MAX_PREV + 1) + MAX_PREV][tid][0]);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(base) : "memory");
#pragma unroll
          for (int jj = 0; jj < MAX_PREV; ++jj)
            if (jj < st.nprev) {
              sa = (unsigned)__cvta_generic_to_shared(&s_epi[f * (MAX_PREV + 1) + jj][tid][0]);
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(st.Fprev[jj][f] + eidx) : "memory");
            }
        }
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    }
    __syncthreads();
    for (int e = tid; e < (TY + 2) * (TXG + 2); e += TXG * TY) {
      const int rr = e / (TXG + 2), gs = e - rr * (TXG + 2);
      const int jj = j0 - 1 + rr, gg = g0 - 1 + gs;
      Vec4<T> vh{0, 0, 0, 0}, vu{0, 0, 0, 0}, vv{0, 0, 0, 0};
      if (!EDGE || (jj >= 0 && jj < Ny && gg >= 0 && gg < ngroups)) {
        const int ib = gg * 4 - OFF;
        const bool plain = !EDGE || !A.apply_bc ||
                           (jj >= (A.ylo ? 1 : 0) && jj <= (A.yhi ? Ny - 3 : Ny - 1) && ib >= 1 && ib + 3 <= Nx - 3);
        if (plain) {
          const size_t o = (size_t)jj * pitch + (size_t)gg * 4;
          vh = ld4(ph + o); vu = ld4(pu + o); vv = ld4(pv + o);
        } else {
          T th[4], tu[4], tv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ii = ib + q;
            th[q] = tu[q] = tv[q] = T(0);
            if (ii >= 0 && ii < Nx) {
              th[q] = swm_bc_value(ph, FH, A.bc, jj, ii, Ny, Nx, pitch, A.ylo, A.yhi, loh, hih);
              tu[q] = swm_bc_value(pu, FU, A.bc, jj, ii, Ny, Nx, pitch, A.ylo, A.yhi, lou, hiu);
              tv[q] = swm_bc_value(pv, FV, A.bc, jj, ii, Ny, Nx, pitch, A.ylo, A.yhi, lov, hiv);
            }
          }
          vh = Vec4<T>{th[0], th[1], th[2], th[3]};
          vu = Vec4<T>{tu[0], tu[1], tu[2], tu[3]};
          vv = Vec4<T>{tv[0], tv[1], tv[2], tv[3]};
        }
      }
      st4(&s_h[rr][gs * 4], vh); st4(&s_u[rr][gs * 4], vu); st4(&s_v[rr][gs * 4], vv);
      Vec4<T> pp = ld4(&s_p[rr][gs * 4]);
      const T gk = A.gprime[k];
      pp.x = pp.x + gk * vh.x; pp.y = pp.y + gk * vh.y; pp.z = pp.z + gk * vh.z; pp.w = pp.w + gk * vh.w;
      st4(&s_p[rr][gs * 4], pp);
    }
    __syncthreads();
    if (EDGE && j >= Ny) continue;     // warp-uniform
    // ---- 3 x 6 register windows: columns i0-1 .. i0+4, rows j-1 .. j+1 ----
    T H[3][6], U[3][6], V[3][6];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const Vec4<T> a = ld4(&s_h[r - 1 + d][cs]);
      const Vec4<T> bq = ld4(&s_u[r - 1 + d][cs]);
      const Vec4<T> c4 = ld4(&s_v[r - 1 + d][cs]);
      T al = __shfl_up_sync(0xffffffffu, a.w, 1), ar = __shfl_down_sync(0xffffffffu, a.x, 1);
      T bl = __shfl_up_sync(0xffffffffu, bq.w, 1), br = __shfl_down_sync(0xffffffffu, bq.x, 1);
      T cl = __shfl_up_sync(0xffffffffu, c4.w, 1), cr = __shfl_down_sync(0xffffffffu, c4.x, 1);
      if (tx == 0) { al = s_h[r - 1 + d][cs - 1]; bl = s_u[r - 1 + d][cs - 1]; cl = s_v[r - 1 + d][cs - 1]; }
      if (tx == TXG - 1) { ar = s_h[r - 1 + d][cs + 4]; br = s_u[r - 1 + d][cs + 4]; cr = s_v[r - 1 + d][cs + 4]; }
      H[d][0] = al; H[d][1] = a.x; H[d][2] = a.y; H[d][3] = a.z; H[d][4] = a.w; H[d][5] = ar;
      U[d][0] = bl; U[d][1] = bq.x; U[d][2] = bq.y; U[d][3] = bq.z; U[d][4] = bq.w; U[d][5] = br;
      V[d][0] = cl; V[d][1] = c4.x; V[d][2] = c4.y; V[d][3] = c4.z; V[d][4] = c4.w; V[d][5] = cr;
    }
    // pressure sum at (j, i..i+4) and (j+1, i..i+3)
    T P0[5], P1[4];
    {
      const Vec4<T> a = ld4(&s_p[r][cs]);
      T ar = __shfl_down_sync(0xffffffffu, a.x, 1);
      if (tx == TXG - 1) ar = s_p[r][cs + 4];
      P0[0] = a.x; P0[1] = a.y; P0[2] = a.z; P0[3] = a.w; P0[4] = ar;
      const Vec4<T> c4 = ld4(&s_p[r + 1][cs]);
      P1[0] = c4.x; P1[1] = c4.y; P1[2] = c4.z; P1[3] = c4.w;
    }
    if (EDGE && g >= ngroups) continue;
    // window index helpers: row d (0..2 <-> dr = d-1), column w (0..5 <-> dc = w-1 from cell 0)
    // ---- potential vorticity at X points: rows dr = -1 (w 1..4) and 0 (w 0..4) ----
    T qm[6], q0[6];
#pragma unroll
    for (int w = 0; w < 5; ++w) {
      q0[w] = T(0); qm[w] = T(0);
      if (rin[1] && cin[w]) {
        const T zeta = (V[1][w + 1] - V[1][w]) * idx_ - (U[2][w] - U[1][w]) * idy_;
        const T hX = T(0.25) * (((H[1][w] + H[1][w + 1]) + H[2][w]) + H[2][w + 1]);
        q0[w] = swm_div(zeta + fX[1], hX);
      }
      if (w >= 1 && rin[0] && cin[w]) {
        const T zeta = (V[0][w + 1] - V[0][w]) * idx_ - (U[1][w] - U[0][w]) * idy_;
        const T hX = T(0.25) * (((H[0][w] + H[0][w + 1]) + H[1][w]) + H[1][w + 1]);
        qm[w] = swm_div(zeta + fX[0], hX);
      }
    }
    // ---- mass fluxes: vh at rows -1,0 (w 1..5); uh at rows 0,1 (w 0..4) ----
    T vhm[6], vh0[6], uh0[6], uh1[6];
#pragma unroll
    for (int w = 0; w < 6; ++w) {
      vhm[w] = (rin[0] && cin[w]) ? (T(0.5) * (H[0][w] + H[1][w])) * V[0][w] : T(0);
      vh0[w] = (rin[1] && cin[w]) ? (T(0.5) * (H[1][w] + H[2][w])) * V[1][w] : T(0);
      if (w < 5) {
        uh0[w] = (rin[1] && cin[w]) ? (T(0.5) * (H[1][w] + H[1][w + 1])) * U[1][w] : T(0);
        uh1[w] = (rin[2] && cin[w]) ? (T(0.5) * (H[2][w] + H[2][w + 1])) * U[2][w] : T(0);
      }
    }
    // ---- kinetic energy: row 0 (w 1..5), row +1 (w 1..4) ----
    T ke0[6], ke1[6];
#pragma unroll
    for (int w = 1; w < 6; ++w) {
      ke0[w] = T(0); ke1[w] = T(0);
      if (rin[1] && cin[w]) {
        const T u2 = T(0.5) * (U[1][w] * U[1][w] + U[1][w - 1] * U[1][w - 1]);
        const T v2 = T(0.5) * (V[1][w] * V[1][w] + V[0][w] * V[0][w]);
        ke0[w] = T(0.5) * (u2 + v2);
      }
      if (w < 5 && rin[2] && cin[w]) {
        const T u2 = T(0.5) * (U[2][w] * U[2][w] + U[2][w - 1] * U[2][w - 1]);
        const T v2 = T(0.5) * (V[2][w] * V[2][w] + V[1][w] * V[1][w]);
        ke1[w] = T(0.5) * (u2 + v2);
      }
    }
    // ---- upwind mass fluxes: fe at row 0 (w 0..4), fn at rows -1, 0 (w 1..4) ----
    T fe[6], fnm[6], fn0[6];
#pragma unroll
    for (int w = 0; w < 5; ++w) {
      fe[w] = T(0); fnm[w] = T(0); fn0[w] = T(0);
      if (rin[1] && cin[w]) {
        const T uu = U[1][w];
        fe[w] = uu * (uu > T(0) ? H[1][w] : H[1][w + 1]);
        const T vv = V[1][w];
        fn0[w] = vv * (vv > T(0) ? H[1][w] : H[2][w]);
      }
      if (rin[0] && cin[w]) {
        const T vv = V[0][w];
        fnm[w] = vv * (vv > T(0) ? H[0][w] : H[1][w]);
      }
    }
    T out_h[4], out_u[4], out_v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = i0 + e, w = e + 1;
      T dh = 0, du = 0, dv = 0;
      if (!EDGE || (i >= 0 && i < Nx)) {
        const bool interior = rin[1] && cin[w];
        if (interior) {
          const T qU = T(0.5) * (q0[w] + qm[w]);
          const T qV = T(0.5) * (q0[w] + q0[w - 1]);
          const T vhU = T(0.25) * (((vh0[w] + vh0[w + 1]) + vhm[w]) + vhm[w + 1]);
          const T uhV = T(0.25) * (((uh0[w] + uh1[w]) + uh0[w - 1]) + uh1[w - 1]);
          const T P00 = ke0[w] + P0[e];
          const T P01 = ke0[w + 1] + P0[e + 1];
          const T P10 = ke1[w] + P1[e];
          du = qU * vhU - (P01 - P00) * idx_;
          dv = -qV * uhV - (P10 - P00) * idy_;
          bool wr = true;
          if (EDGE && (A.spec & SOMAX_B200_SPEC_ADVECTION_REGION2))
            wr = (j >= (A.ylo ? 2 : 0) && j <= (A.yhi ? Ny - 3 : Ny - 1) && i >= 2 && i <= Nx - 3);
          if (wr) dh = -((fe[w] - fe[w - 1]) * idx_ + (fn0[w] - fnm[w]) * idy_);
        }
        if (k == 0) { du = du + windx; dv = dv + windy; }
        if (interior) {
          T lu, lv;
          if (A.spec & SOMAX_B200_SPEC_DIFFUSION_FLUX) {
            auto fxf = [&](const T (&X)[3][6], int d, int ww) -> T {
              return (rin[d] && cin[ww]) ? (X[d][ww + 1] - X[d][ww]) * nux : T(0); };
            auto fyf = [&](const T (&X)[3][6], int d, int ww) -> T {
              return (rin[d] && cin[ww]) ? (X[d + 1][ww] - X[d][ww]) * nuy : T(0); };
            lu = (fxf(U, 1, w) - fxf(U, 1, w - 1)) * idx_ + (fyf(U, 1, w) - fyf(U, 0, w)) * idy_;
            lv = (fxf(V, 1, w) - fxf(V, 1, w - 1)) * idx_ + (fyf(V, 1, w) - fyf(V, 0, w)) * idy_;
          } else {
            lu = A.nu * ((U[1][w + 1] - T(2) * U[1][w] + U[1][w - 1]) * A.idx2 +
                         (U[2][w] - T(2) * U[1][w] + U[0][w]) * A.idy2);
            lv = A.nu * ((V[1][w + 1] - T(2) * V[1][w] + V[1][w - 1]) * A.idx2 +
                         (V[2][w] - T(2) * V[1][w] + V[0][w]) * A.idy2);
          }
          du = du + lu; dv = dv + lv;
        }
        if (k == L.nl - 1) { du = du + (-A.kappa * U[1][w]); dv = dv + (-A.kappa * V[1][w]); }
      }
      out_h[e] = dh; out_u[e] = du; out_v[e] = dv;
    }
    const size_t idx = plane_off + (size_t)j * pitch + (size_t)g * 4;
    if (!STAGE_EPI) {
      rk_epilogue4_fast(st, FH, idx, Vec4<T>{out_h[0], out_h[1], out_h[2], out_h[3]});
      rk_epilogue4_fast(st, FU, idx, Vec4<T>{out_u[0], out_u[1], out_u[2], out_u[3]});
      rk_epilogue4_fast(st, FV, idx, Vec4<T>{out_v[0], out_v[1], out_v[2], out_v[3]});
    } else {
      asm volatile("cp.async.wait_group 0;\n" ::: "memory");
      const Vec4<T> Fv[3] = {Vec4<T>{out_h[0], out_h[1], out_h[2], out_h[3]},
                             Vec4<T>{out_u[0], out_u[1], out_u[2], out_u[3]},
                             Vec4<T>{out_v[0], out_v[1], out_v[2], out_v[3]}};
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        const Vec4<T> F = Fv[f];
        if (st.Fout[f]) st4(st.Fout[f] + idx, F);
        if (!st.Yout[f]) continue;
        Vec4<T> acc = ld4(&s_epi[f * (MAX_PREV + 1) + MAX_PREV][tid][0]);
#pragma unroll
        for (int jj = 0; jj < MAX_PREV; ++jj) {
          if (jj < st.nprev) {
            const Vec4<T> kk = ld4(&s_epi[f * (MAX_PREV + 1) + jj][tid][0]);
            acc.x = fma(st.adt[jj], kk.x, acc.x); acc.y = fma(st.adt[jj], kk.y, acc.y);
            acc.z = fma(st.adt[jj], kk.z, acc.z); acc.w = fma(st.adt[jj], kk.w, acc.w);
          }
        }
        acc.x = fma(st.adt_new, F.x, acc.x); acc.y = fma(st.adt_new, F.y, acc.y);
        acc.z = fma(st.adt_new, F.z, acc.z); acc.w = fma(st.adt_new, F.w, acc.w);
        st4(st.Yout[f] + idx, acc);
      }
    }
  }
}

// Interior CTAs (97 % at 4096^2): no predicate, and a software pipeline over the CTA's work items
// (tile t of the CTA's column of `ntile` row tiles, layer k): the h/u/v halo tile of item it+1 is
// fetched with 16-byte cp.async into the other half of a double buffer while item `it` is computed,
// next to the Tsit5 epilogue operands of item `it` (per-thread slots).  Commit groups retire in
// order, so "all but the newest group" = tile `it` has landed.  The layer pressure sum lives in
// registers (the general path keeps a fourth shared tile for it).  Before: tile load -> barrier ->
// stencil per layer, the load latency covered only by the SM's other CTA.
// NPREV (stored stage derivatives entering the Runge-Kutta combination) and FLUX (flux-form
// diffusion) are compile-time: the kernel is issue bound, and the `jj < nprev` / `spec &` tests of
// the generic path were a tenth of its instructions.
template <typename T, int NPREV, bool FLUX>
__device__ __forceinline__ void swm_rhs_inner(const SwmArgs<T>& A, const Stage<T>& st, int jt0, int ntile,
                                              T* __restrict__ tiles, T (*s_epi)[TXG * TY][4]) {
  constexpr int TSZ = (TY + 2) * SSW;            // one field's tile
  const Layout& L = A.L;
  const int pitch = L.pitch, nl = L.nl;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TXG + tx;
  const int g0 = blockIdx.x * TXG, g = g0 + tx, b = blockIdx.z;
  const int r = ty + 1, cs = 4 * (tx + 1);
  const T idx_ = A.idx, idy_ = A.idy;
  const T nux = A.nu * idx_, nuy = A.nu * idy_;
  const int nitems = ntile * nl;
  const bool want_y = st.Yout[FH] != nullptr;

  // the (TY+2) x (TXG+2) vectors of a tile over the 256 threads: elements tid and tid + 256
  // (the second one for the first 84 threads); offsets relative to the tile's first vector
  constexpr int NE = (TY + 2) * (TXG + 2);
  const int e1 = tid + TXG * TY;
  const int rr0 = tid / (TXG + 2), gs0 = tid - rr0 * (TXG + 2);
  const int rr1 = e1 / (TXG + 2), gs1 = e1 - rr1 * (TXG + 2);
  const unsigned goff0 = (unsigned)(rr0 * pitch + gs0 * 4), goff1 = (unsigned)(rr1 * pitch + gs1 * 4);
  const unsigned tiles_sa = (unsigned)__cvta_generic_to_shared(tiles);
  const unsigned soff0 = (unsigned)((rr0 * SSW + gs0 * 4) * sizeof(T)), soff1 = (unsigned)((rr1 * SSW + gs1 * 4) * sizeof(T));
  const unsigned epi_sa = (unsigned)__cvta_generic_to_shared(&s_epi[0][tid][0]);
  constexpr unsigned EPI_STRIDE = TXG * TY * 4 * sizeof(T);

  auto issue_tile = [&](int it, size_t base) {
    const unsigned sb = tiles_sa + (unsigned)((it & 1) * 3 * TSZ * sizeof(T));
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const T* src = st.Yin[f] + base;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sb + soff0 + f * TSZ * (unsigned)sizeof(T)), "l"(src + goff0) : "memory");
      if (e1 < NE)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sb + soff1 + f * TSZ * (unsigned)sizeof(T)), "l"(src + goff1) : "memory");
    }
  };
  auto issue_epi = [&](size_t eidx) {
    if (!want_y) return;
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const T* base = (st.y[f] ? st.y[f] : st.Yin[f]) + eidx;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(epi_sa + (f * (MAX_PREV + 1) + MAX_PREV) * EPI_STRIDE), "l"(base) : "memory");
#pragma unroll
      for (int jj = 0; jj < NPREV; ++jj)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(epi_sa + (f * (MAX_PREV + 1) + jj) * EPI_STRIDE), "l"(st.Fprev[jj][f] + eidx) : "memory");
    }
  };

  // (tile, layer) of the item, its tile base and this thread's output vector, advanced incrementally
  const size_t plane = L.plane();
  const size_t tile_step = (size_t)TY * pitch - (size_t)(nl - 1) * plane;      // last layer -> layer 0 of the next tile
  size_t tb = (size_t)b * nl * plane + (size_t)(jt0 - 1) * pitch + (size_t)(g0 - 1) * 4;
  size_t idx = (size_t)b * nl * plane + (size_t)(jt0 + ty) * pitch + (size_t)g * 4;
  int k = 0, j = jt0 + ty;
  issue_tile(0, tb);
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  T P0[5], P1[4], fX[2], windx = 0, windy = 0;
  for (int it = 0; it < nitems; ++it) {
    const bool last_layer = k == nl - 1;
    const size_t tbn = tb + (last_layer ? tile_step : plane);
    // row constants of a new tile: loaded here, used after the barrier that hides their latency
    T fr[3] = {0, 0, 0}, wxr = 0, wyr = 0;
    if (k == 0) { fr[0] = A.f[j - 1]; fr[1] = A.f[j]; fr[2] = A.f[j + 1]; wxr = A.wx[j]; wyr = A.wy[j]; }
    issue_epi(idx);
    if (it + 1 < nitems) issue_tile(it + 1, tbn);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
    if (k == 0) {
      // Coriolis at the X points of rows j-1 and j (f depends on y only; reference summation order)
#pragma unroll
      for (int d = 0; d < 2; ++d) fX[d] = T(0.25) * (((fr[d] + fr[d]) + fr[d + 1]) + fr[d + 1]);
      windx = (A.tau0 * wxr) * A.iH0; windy = (A.tau0 * wyr) * A.iH0;
#pragma unroll
      for (int e = 0; e < 5; ++e) P0[e] = T(0);
#pragma unroll
      for (int e = 0; e < 4; ++e) P1[e] = T(0);
    }
    // ---- 3 x 6 register windows: columns i0-1 .. i0+4, rows j-1 .. j+1 ----
    const T* bh = tiles + (it & 1) * 3 * TSZ;
    T H[3][6], U[3][6], V[3][6];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const T* rowp = bh + (r - 1 + d) * SSW + cs;
      const Vec4<T> a = ld4(rowp);
      const Vec4<T> bq = ld4(rowp + TSZ);
      const Vec4<T> c4 = ld4(rowp + 2 * TSZ);
      T al = __shfl_up_sync(0xffffffffu, a.w, 1), ar = __shfl_down_sync(0xffffffffu, a.x, 1);
      T bl = __shfl_up_sync(0xffffffffu, bq.w, 1), br = __shfl_down_sync(0xffffffffu, bq.x, 1);
      T cl = __shfl_up_sync(0xffffffffu, c4.w, 1), cr = __shfl_down_sync(0xffffffffu, c4.x, 1);
      if (tx == 0) { al = rowp[-1]; bl = rowp[TSZ - 1]; cl = rowp[2 * TSZ - 1]; }
      if (tx == TXG - 1) { ar = rowp[4]; br = rowp[TSZ + 4]; cr = rowp[2 * TSZ + 4]; }
      H[d][0] = al; H[d][1] = a.x; H[d][2] = a.y; H[d][3] = a.z; H[d][4] = a.w; H[d][5] = ar;
      U[d][0] = bl; U[d][1] = bq.x; U[d][2] = bq.y; U[d][3] = bq.z; U[d][4] = bq.w; U[d][5] = br;
      V[d][0] = cl; V[d][1] = c4.x; V[d][2] = c4.y; V[d][3] = c4.z; V[d][4] = c4.w; V[d][5] = cr;
    }
    __syncthreads();      // the buffer may be refilled (by the tile issued at the top of the next item)
    // pressure sum at (j, i..i+4) and (j+1, i..i+3)
    {
      const T gk = A.gprime[k];
#pragma unroll
      for (int e = 0; e < 5; ++e) P0[e] = P0[e] + gk * H[1][1 + e];
#pragma unroll
      for (int e = 0; e < 4; ++e) P1[e] = P1[e] + gk * H[2][1 + e];
    }
    // ---- potential vorticity at X points: rows dr = -1 (w 1..4) and 0 (w 0..4) ----
    T qm[6], q0[6];
#pragma unroll
    for (int w = 0; w < 5; ++w) {
      {
        const T zeta = (V[1][w + 1] - V[1][w]) * idx_ - (U[2][w] - U[1][w]) * idy_;
        const T hX = T(0.25) * (((H[1][w] + H[1][w + 1]) + H[2][w]) + H[2][w + 1]);
        q0[w] = swm_div(zeta + fX[1], hX);
      }
      if (w >= 1) {
        const T zeta = (V[0][w + 1] - V[0][w]) * idx_ - (U[1][w] - U[0][w]) * idy_;
        const T hX = T(0.25) * (((H[0][w] + H[0][w + 1]) + H[1][w]) + H[1][w + 1]);
        qm[w] = swm_div(zeta + fX[0], hX);
      }
    }
    // Factors 1/2 and 1/4 are exact in binary floating point, so they are collected: twice the mass
    // fluxes, four times the kinetic energy and twice the interpolated q are carried, and one
    // power-of-two constant per use restores the scale - bit-identical to scaling every term.
    // ---- 2 x mass fluxes: vh at rows -1,0 (w 1..5); uh at rows 0,1 (w 0..4) ----
    T vhm[6], vh0[6], uh0[6], uh1[6];
#pragma unroll
    for (int w = 0; w < 6; ++w) {
      vhm[w] = (H[0][w] + H[1][w]) * V[0][w];
      vh0[w] = (H[1][w] + H[2][w]) * V[1][w];
      if (w < 5) {
        uh0[w] = (H[1][w] + H[1][w + 1]) * U[1][w];
        uh1[w] = (H[2][w] + H[2][w + 1]) * U[2][w];
      }
    }
    // ---- 4 x kinetic energy: row 0 (w 1..5), row +1 (w 1..4) ----
    T ke0[6], ke1[6];
#pragma unroll
    for (int w = 1; w < 6; ++w) {
      {
        const T u2 = U[1][w] * U[1][w] + U[1][w - 1] * U[1][w - 1];
        const T v2 = V[1][w] * V[1][w] + V[0][w] * V[0][w];
        ke0[w] = u2 + v2;
      }
      if (w < 5) {
        const T u2 = U[2][w] * U[2][w] + U[2][w - 1] * U[2][w - 1];
        const T v2 = V[2][w] * V[2][w] + V[1][w] * V[1][w];
        ke1[w] = u2 + v2;
      }
    }
    // ---- upwind mass fluxes: fe at row 0 (w 0..4), fn at rows -1, 0 (w 1..4) ----
    T fe[6], fnm[6], fn0[6];
#pragma unroll
    for (int w = 0; w < 5; ++w) {
      const T uu = U[1][w];
      fe[w] = uu * (uu > T(0) ? H[1][w] : H[1][w + 1]);
      const T vv = V[1][w];
      fn0[w] = vv * (vv > T(0) ? H[1][w] : H[2][w]);
      const T vm = V[0][w];
      fnm[w] = vm * (vm > T(0) ? H[0][w] : H[1][w]);
    }
    Vec4<T> Fv[3];
    {
      T out_h[4], out_u[4], out_v[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int w = e + 1;
        const T qU = q0[w] + qm[w];                                   // 2 qU
        const T qV = q0[w] + q0[w - 1];                               // 2 qV
        const T vhU = T(0.0625) * (((vh0[w] + vh0[w + 1]) + vhm[w]) + vhm[w + 1]);      // vhU / 2
        const T uhV = T(0.0625) * (((uh0[w] + uh1[w]) + uh0[w - 1]) + uh1[w - 1]);      // uhV / 2
        const T P00 = fma(T(0.25), ke0[w], P0[e]);
        const T P01 = fma(T(0.25), ke0[w + 1], P0[e + 1]);
        const T P10 = fma(T(0.25), ke1[w], P1[e]);
        T du = qU * vhU - (P01 - P00) * idx_;
        T dv = -qV * uhV - (P10 - P00) * idy_;
        const T dh = -((fe[w] - fe[w - 1]) * idx_ + (fn0[w] - fnm[w]) * idy_);
        if (k == 0) { du = du + windx; dv = dv + windy; }
        T lu, lv;
        if (FLUX) {
          // (nu / dx rounded once: <= 1 ulp from nu * (d / dx), like the other constant divisors)
          auto fxf = [&](const T (&X)[3][6], int d, int ww) -> T { return (X[d][ww + 1] - X[d][ww]) * nux; };
          auto fyf = [&](const T (&X)[3][6], int d, int ww) -> T { return (X[d + 1][ww] - X[d][ww]) * nuy; };
          lu = (fxf(U, 1, w) - fxf(U, 1, w - 1)) * idx_ + (fyf(U, 1, w) - fyf(U, 0, w)) * idy_;
          lv = (fxf(V, 1, w) - fxf(V, 1, w - 1)) * idx_ + (fyf(V, 1, w) - fyf(V, 0, w)) * idy_;
        } else {
          lu = A.nu * ((U[1][w + 1] - T(2) * U[1][w] + U[1][w - 1]) * A.idx2 +
                       (U[2][w] - T(2) * U[1][w] + U[0][w]) * A.idy2);
          lv = A.nu * ((V[1][w + 1] - T(2) * V[1][w] + V[1][w - 1]) * A.idx2 +
                       (V[2][w] - T(2) * V[1][w] + V[0][w]) * A.idy2);
        }
        du = du + lu; dv = dv + lv;
        if (last_layer) { du = du + (-A.kappa * U[1][w]); dv = dv + (-A.kappa * V[1][w]); }
        out_h[e] = dh; out_u[e] = du; out_v[e] = dv;
      }
      Fv[0] = Vec4<T>{out_h[0], out_h[1], out_h[2], out_h[3]};
      Fv[1] = Vec4<T>{out_u[0], out_u[1], out_u[2], out_u[3]};
      Fv[2] = Vec4<T>{out_v[0], out_v[1], out_v[2], out_v[3]};
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#pragma unroll
    for (int f = 0; f < 3; ++f) {
      const Vec4<T> F = Fv[f];
      if (st.Fout[f]) st4(st.Fout[f] + idx, F);
      if (!st.Yout[f]) continue;
      Vec4<T> acc = ld4(&s_epi[f * (MAX_PREV + 1) + MAX_PREV][tid][0]);
#pragma unroll
      for (int jj = 0; jj < NPREV; ++jj) {
        const Vec4<T> kk = ld4(&s_epi[f * (MAX_PREV + 1) + jj][tid][0]);
        acc.x = fma(st.adt[jj], kk.x, acc.x); acc.y = fma(st.adt[jj], kk.y, acc.y);
        acc.z = fma(st.adt[jj], kk.z, acc.z); acc.w = fma(st.adt[jj], kk.w, acc.w);
      }
      acc.x = fma(st.adt_new, F.x, acc.x); acc.y = fma(st.adt_new, F.y, acc.y);
      acc.z = fma(st.adt_new, F.z, acc.z); acc.w = fma(st.adt_new, F.w, acc.w);
      st4(st.Yout[f] + idx, acc);
    }
    tb = tbn;
    idx += last_layer ? tile_step : plane;
    j += last_layer ? TY : 0;
    k = last_layer ? 0 : k + 1;
  }
}

template <typename T, int NPREV, bool FLUX>
__global__ void __launch_bounds__(TXG* TY, (sizeof(T) == 4 ? 2 : 1))
swm_rhs_kernel_fast(SwmArgs<T> A, Stage<T> st, int ntile) {
  constexpr int TSZ = (TY + 2) * SSW;
  __shared__ __align__(32) T tiles[2 * 3 * TSZ];      // inner: double-buffered h, u, v; edge: h, u, v, p
  extern __shared__ __align__(16) unsigned char swm_dyn[];
  // A tile is `inner` when its rows j0-1 .. j0+TY and column groups g0-1 .. g0+TXG (halo included)
  // are all plain loads (two cells inside a physical boundary, which also covers the [2:-2] write
  // region of the advection) and all in range.  Runs of inner tiles go through the pipeline.
  const int Ny = A.L.Ny, Nx = A.L.Nx;
  const int jt0 = (int)blockIdx.y * ntile * TY;
  const int ilo = ((int)blockIdx.x * TXG - 1) * 4 - OFF, ihi = ((int)blockIdx.x * TXG + TXG) * 4 - OFF + 3;
  const bool cols_inner = ilo >= 1 && ihi <= Nx - 3 && (int)blockIdx.x * TXG + TXG < A.L.groups();
  auto tile_inner = [&](int t) {
    const int j0 = jt0 + t * TY;
    return cols_inner && j0 - 1 >= (A.ylo ? 1 : 0) && j0 + TY <= (A.yhi ? Ny - 3 : Ny - 1);
  };
  T (*s_h)[SSW] = reinterpret_cast<T (*)[SSW]>(tiles);
  for (int t = 0; t < ntile;) {
    const int j0 = jt0 + t * TY;
    if (j0 >= Ny) break;
    if (t) __syncthreads();      // the previous tile's windows have been read
    if (tile_inner(t)) {
      int n = 1;
      while (t + n < ntile && tile_inner(t + n)) ++n;
      swm_rhs_inner<T, NPREV, FLUX>(A, st, j0, n, tiles, reinterpret_cast<T (*)[TXG * TY][4]>(swm_dyn));
      t += n;
    } else {
      swm_rhs_body<T, true>(A, st, j0, s_h, s_h + (TY + 2), s_h + 2 * (TY + 2), s_h + 3 * (TY + 2));
      ++t;
    }
  }
}

// In-place apply_boundary_conditions on padded planes.
template <typename T>
__global__ void swm_bc_kernel(T* __restrict__ h, T* __restrict__ u, T* __restrict__ v, Layout L,
                              int bc, int ylo, int yhi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int j = blockIdx.y;
  int p = blockIdx.z;
  if (i >= L.Nx) return;
  bool ring = ((j == 0 && ylo) || (j == L.Ny - 1 && yhi) || i == 0 || i == L.Nx - 1);
  bool near_u = (bc == SOMAX_B200_BC_WALL) && (i == L.Nx - 2);
  bool near_v = (bc == SOMAX_B200_BC_WALL) && yhi && (j == L.Ny - 2);
  if (!ring && !near_u && !near_v) return;
  size_t po = (size_t)p * L.plane();
  size_t o = po + (size_t)j * L.pitch + OFF + i;
  // every value is computed from cells the kernel never writes (see swm_bc_value)
  if (ring) h[o] = swm_bc_value(h + po, FH, bc, j, i, L.Ny, L.Nx, L.pitch, ylo, yhi);
  if (ring || near_u) u[o] = swm_bc_value(u + po, FU, bc, j, i, L.Ny, L.Nx, L.pitch, ylo, yhi);
  if (ring || near_v) v[o] = swm_bc_value(v + po, FV, bc, j, i, L.Ny, L.Nx, L.pitch, ylo, yhi);
}

// Diagnostics on the reference layout (state as given, no BC): swm/multilayer.py:225-256.
template <typename T>
__global__ void swm_diag_kernel(const T* __restrict__ h, const T* __restrict__ u,
                                const T* __restrict__ v, const T* __restrict__ f, int f_cp,
                                int f_xs, int nl, int Ny, int Nx, double dx, double dy,
                                double* __restrict__ out) {
  const int b = blockIdx.z / nl, k = blockIdx.z % nl;
  const size_t po = ((size_t)b * nl + k) * Ny * Nx;
  const T* H = h + po; const T* U = u + po; const T* V = v + po;
  double ke = 0, h2 = 0, ens = 0, bad = 0;
  for (int j = blockIdx.y; j < Ny; j += gridDim.y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Nx; i += gridDim.x * blockDim.x) {
      size_t o = (size_t)j * Nx + i;
      T hv = H[o], uv = U[o], vv = V[o];
      if (!isfinite((double)hv)) bad += 1;
      if (!isfinite((double)uv)) bad += 1;
      if (!isfinite((double)vv)) bad += 1;
      if (j >= 1 && j <= Ny - 2 && i >= 1 && i <= Nx - 2) {
        T uw = U[o - 1], vs = V[o - Nx];
        T kev = T(0.5) * (T(0.5) * (uv * uv + uw * uw) + T(0.5) * (vv * vv + vs * vs));
        T zeta = (T)((V[o + 1] - vv) / (T)dx - (U[o + Nx] - uv) / (T)dy);
        auto Fc = [&](int jj, int ii) { return f[(size_t)jj * f_cp + (size_t)ii * f_xs]; };
        T fX = T(0.25) * (Fc(j, i) + Fc(j, i + 1) + Fc(j + 1, i) + Fc(j + 1, i + 1));
        T hX = T(0.25) * (hv + H[o + 1] + H[o + Nx] + H[o + Nx + 1]);
        T q = (zeta + fX) / hX;
        ke += (double)kev;
        h2 += (double)(hv * hv);
        ens += (double)(q * q * hX);
      }
    }
  }
  // block reduce
  __shared__ double red[4][8];
  double vals[4] = {ke, h2, ens, bad};
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    double x = vals[q];
    for (int s = 16; s > 0; s >>= 1) x += __shfl_down_sync(0xffffffffu, x, s);
    if (lane == 0) red[q][w] = x;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double x = 0;
    for (int ww = 0; ww < (int)(blockDim.x >> 5); ++ww) x += red[threadIdx.x][ww];
    double area = dx * dy;
    double* o = out + (size_t)b * (3 * nl + 1);
    if (threadIdx.x == 0) atomicAdd(o + k, x * area);
    if (threadIdx.x == 1) atomicAdd(o + nl + k, x * area);
    if (threadIdx.x == 2) atomicAdd(o + 2 * nl + k, 0.5 * x * area);
    if (threadIdx.x == 3) atomicAdd(o + 3 * nl, x);
  }
}

}  // namespace sb

using namespace sb;

struct somax_b200_swm_s {
  int dtype;
  Layout L;
  int ny, nx, bc;
  unsigned spec;
  double dx, dy;
  double gprime[SWM_MAX_NL];
  void* f = nullptr; void* wx = nullptr; void* wy = nullptr;
  bool f1d = false, wx1d = false, wy1d = false;
  void* y[3] = {0, 0, 0};
  void* Ya[3] = {0, 0, 0};
  void* Yb[3] = {0, 0, 0};
  void* F[5][3] = {};
  StepGraph graph;
  size_t bytes = 0;
  int bc_ylo = 1, bc_yhi = 1;   // physical ghost rows (a slab of the distributed model clears them)
  // geostrophic projection (reparameterized QG): Helmholtz solver, PV / streamfunction planes and the
  // projected state the right-hand side is evaluated at
  QgSolver* proj = nullptr;
  ProjArgs pargs;
  void* pq = nullptr; void* ppsi = nullptr; void* P[3] = {0, 0, 0};
};

namespace {

template <typename T>
SwmArgs<T> make_args(somax_b200_swm_t h, const somax_b200_params* p, int apply_bc) {
  SwmArgs<T> A;
  A.L = h->L; A.bc = h->bc; A.spec = h->spec; A.apply_bc = apply_bc; A.ylo = h->bc_ylo; A.yhi = h->bc_yhi;
  for (int f = 0; f < 3; ++f) { A.lo_src[f] = nullptr; A.hi_src[f] = nullptr; }
  A.dx = (T)h->dx; A.dy = (T)h->dy; A.dx2 = (T)(h->dx * h->dx); A.dy2 = (T)(h->dy * h->dy);
  A.idx = (T)(1.0 / h->dx); A.idy = (T)(1.0 / h->dy); A.idx2 = (T)(1.0 / (h->dx * h->dx));
  A.idy2 = (T)(1.0 / (h->dy * h->dy)); A.iH0 = (T)(1.0 / p->H0);
  A.f = (const T*)h->f;   A.f_cp = h->f1d ? 1 : h->L.Nx;   A.f_xs = h->f1d ? 0 : 1;
  A.wx = (const T*)h->wx; A.wx_cp = h->wx1d ? 1 : h->L.Nx; A.wx_xs = h->wx1d ? 0 : 1;
  A.wy = (const T*)h->wy; A.wy_cp = h->wy1d ? 1 : h->L.Nx; A.wy_xs = h->wy1d ? 0 : 1;
  for (int k = 0; k < SWM_MAX_NL; ++k) A.gprime[k] = (T)h->gprime[k];
  A.H0 = (T)p->H0; A.nu = (T)p->lateral_viscosity; A.kappa = (T)p->bottom_drag;
  A.tau0 = (T)p->wind_amplitude;
  return A;
}

// out = project(BC(in)) (apply_bc) or project(in); in / out are triples of padded planes.
template <typename T>
int project_state(somax_b200_swm_t h, const T* const in[3], T* const out[3], int apply_bc, cudaStream_t s) {
  const Layout& L = h->L;
  dim3 b(128), g((L.Nx + 127) / 128, L.Ny, L.batch * L.nl);
  prof_begin("rqg_q_kernel", s);
  rqg_q_kernel<T><<<g, b, 0, s>>>(in[0], in[1], in[2], (T*)h->pq, L, h->bc, apply_bc, h->pargs, (T)h->dx, (T)h->dy);
  SB_LAUNCH_CHECK();
  // the PV carries its ring (-f0 eta / H there): the solve takes it as it is, psi's ring is zeroed
  if (int rc = qg_solver_run<T>(h->proj, (const T*)h->pq, (T*)h->ppsi, 0, 0, s)) return rc;
  prof_begin("rqg_g_kernel", s);
  rqg_g_kernel<T><<<g, b, 0, s>>>((const T*)h->ppsi, out[0], out[1], out[2], L, h->pargs, (T)h->dx, (T)h->dy);
  SB_LAUNCH_CHECK();
  return 0;
}

template <typename T, int NPREV, bool FLUX>
static int launch_fast_one(const SwmArgs<T>& A, const Stage<T>& st, int ntile, dim3 grid, dim3 block, size_t dyn,
                           cudaStream_t s) {
  if (int rc = ensure_dyn_smem((const void*)swm_rhs_kernel_fast<T, NPREV, FLUX>, dyn)) return rc;
  swm_rhs_kernel_fast<T, NPREV, FLUX><<<grid, block, dyn, s>>>(A, st, ntile);
  return 0;
}

template <typename T>
static int launch_fast(const SwmArgs<T>& A, const Stage<T>& st, int ntile, dim3 grid, dim3 block, size_t dyn,
                       cudaStream_t s) {
  const bool flux = (A.spec & SOMAX_B200_SPEC_DIFFUSION_FLUX) != 0;
#define SB_SWM_CASE(N)                                                                        \
  case N:                                                                                     \
    return flux ? launch_fast_one<T, N, true>(A, st, ntile, grid, block, dyn, s)              \
                : launch_fast_one<T, N, false>(A, st, ntile, grid, block, dyn, s);
  switch (st.nprev) {
    SB_SWM_CASE(0) SB_SWM_CASE(1) SB_SWM_CASE(2) SB_SWM_CASE(3) SB_SWM_CASE(4) SB_SWM_CASE(5)
  }
#undef SB_SWM_CASE
  return SOMAX_B200_ERR_INVALID;
}

template <typename T>
int launch_rhs(somax_b200_swm_t h, const SwmArgs<T>& A_in, const Stage<T>& st_in, cudaStream_t s) {
  const Layout& L = h->L;
  Stage<T> st = st_in;
  SwmArgs<T> A = A_in;
  if (h->proj && A.apply_bc) {
    // ReparameterizedQG._rhs = swm.vector_field(project(swm_bc(Y))): the right-hand side is evaluated
    // at the projected state, the Runge-Kutta combination still starts from the un-projected one
    const T* in[3] = {st.Yin[0], st.Yin[1], st.Yin[2]};
    T* out[3] = {(T*)h->P[0], (T*)h->P[1], (T*)h->P[2]};
    if (int rc = project_state<T>(h, in, out, 1, s)) return rc;
    for (int f = 0; f < 3; ++f) {
      if (!st.y[f]) st.y[f] = st.Yin[f];
      st.Yin[f] = (const T*)h->P[f];
    }
    A.apply_bc = 0;
  }
  stage_finalize(st, (double)st.dt);
  dim3 block(TXG, TY);
  dim3 grid((L.groups() + TXG - 1) / TXG, (L.Ny + TY - 1) / TY, L.batch);
  // fast kernel: a CTA can walk `ntile` row tiles (software pipeline over tiles and layers).  One
  // tile per CTA - the pipeline then only spans the layers - measured best at 2 x 4096^2: 2.91 ms/step
  // against 2.97 / 3.04 / 3.11 / 3.44 for 2 / 3 / 4 / 8 tiles (more, shorter CTAs balance the SMs
  // better than a longer pipeline hides latency); SOMAX_B200_SWM_NTILE overrides for tuning.
  int ntile = 1;
  if (const char* e = getenv("SOMAX_B200_SWM_NTILE")) ntile = std::max(1, atoi(e));
  prof_begin("swm_rhs_kernel", s);
  if constexpr (sizeof(T) == 8) {
    // fp64 = the validation pipeline: reference operation order, no FMA contraction (swm_f64.cu)
    if (int rc = swm_launch_reference_order_f64(A, st, grid, block, s)) return rc;
  } else if (h->f1d && h->wx1d && h->wy1d) {
    // fp32: per-thread shared slots for the prefetched epilogue operands
    const size_t dyn = sizeof(T) == 4 ? (size_t)3 * (MAX_PREV + 1) * TXG * TY * 4 * sizeof(T) : 0;
    dim3 gridf(grid.x, (grid.y + ntile - 1) / ntile, grid.z);
    if (int rc = launch_fast<T>(A, st, ntile, gridf, block, dyn, s)) return rc;
  }
  else swm_rhs_kernel<T><<<grid, block, 0, s>>>(A, st);
  SB_LAUNCH_CHECK();
  return 0;
}

template <typename T>
Stage<T> empty_stage() {
  Stage<T> st;
  st.nfields = 3; st.nprev = 0;
  for (int f = 0; f < MAX_FIELDS; ++f) {
    st.Yin[f] = nullptr; st.y[f] = nullptr; st.Fout[f] = nullptr; st.Yout[f] = nullptr;
    for (int j = 0; j < MAX_PREV; ++j) st.Fprev[j][f] = nullptr;
  }
  for (int j = 0; j < MAX_PREV; ++j) st.a[j] = 0;
  st.a_new = 0; st.dt = 0;
  return st;
}

template <typename T>
int bc_inplace(somax_b200_swm_t h, void* const f[3], cudaStream_t s) {
  const Layout& L = h->L;
  dim3 b(256), g((L.Nx + 255) / 256, L.Ny, L.batch * L.nl);
  prof_begin("swm_bc_kernel", s);
  swm_bc_kernel<T><<<g, b, 0, s>>>((T*)f[0], (T*)f[1], (T*)f[2], L, h->bc, h->bc_ylo, h->bc_yhi);
  SB_LAUNCH_CHECK();
  if (h->proj) {
    // ReparameterizedQG.apply_boundary_conditions: the shallow-water BCs, then the projection
    const T* in[3] = {(const T*)f[0], (const T*)f[1], (const T*)f[2]};
    T* out[3] = {(T*)h->P[0], (T*)h->P[1], (T*)h->P[2]};
    if (int rc = project_state<T>(h, in, out, 0, s)) return rc;
    const size_t nb = L.count() * sizeof(T);
    for (int k = 0; k < 3; ++k) SB_CUDA(cudaMemcpyAsync(f[k], h->P[k], nb, cudaMemcpyDeviceToDevice, s));
  }
  return 0;
}

template <typename T>
int swm_steps_impl(somax_b200_swm_t h, void* hh, void* u, void* v, long n_steps, double dt,
                   double dt_last, const somax_b200_params* p, bool bc0, cudaStream_t caller) {
  const Layout& L = h->L;
  const long total = n_steps + (dt_last > 0 ? 1 : 0);
  const bool use_graph = !prof_enabled() && L.count() <= GRAPH_MAX_CELLS && n_steps >= 9 && !h->proj &&
                         h->graph.init() == 0;
  cudaStream_t s = caller;
  if (use_graph) {
    s = h->graph.stream;
    SB_CUDA(cudaEventRecord(h->graph.ev_in, caller));
    SB_CUDA(cudaStreamWaitEvent(s, h->graph.ev_in, 0));
  }
  void* ext[3] = {hh, u, v};
  void *y[3], *Yc[3], *Yn[3];
  for (int f = 0; f < 3; ++f) { y[f] = h->y[f]; Yc[f] = h->Ya[f]; Yn[f] = h->Yb[f]; }
  for (int f = 0; f < 3; ++f)
    if (int rc = pack_field<T>((const T*)ext[f], (T*)y[f], L, s)) return rc;
  if (bc0)
    if (int rc = bc_inplace<T>(h, y, s)) return rc;   // integrate(): BC on state0 (not when resuming)
  if (total > 0) {
    SwmArgs<T> A = make_args<T>(h, p, 1);
    auto step_dt = [&](long i) { return (i < n_steps) ? dt : dt_last; };
    // k1 of the first step: F1 = f(BC(y)), Y2 = y + a21*dt*F1
    {
      Stage<T> st = empty_stage<T>();
      for (int f = 0; f < 3; ++f) {
        st.Yin[f] = (const T*)y[f]; st.Fout[f] = (T*)h->F[0][f]; st.Yout[f] = (T*)Yc[f];
      }
      st.a_new = (T)TSIT5_A[0][0]; st.dt = (T)step_dt(0);
      if (int rc = launch_rhs<T>(h, A, st, s)) return rc;
    }
    auto stages = [&](double hd) -> int {
      for (int e = 1; e <= 5; ++e) {
        Stage<T> st = empty_stage<T>();
        st.nprev = e; st.dt = (T)hd; st.a_new = (T)TSIT5_A[e][e];
        for (int jj = 0; jj < e; ++jj) st.a[jj] = (T)TSIT5_A[e][jj];
        for (int f = 0; f < 3; ++f) {
          st.Yin[f] = (const T*)Yc[f]; st.y[f] = (const T*)y[f]; st.Yout[f] = (T*)Yn[f];
          st.Fout[f] = (e <= 4) ? (T*)h->F[e][f] : nullptr;
          for (int jj = 0; jj < e; ++jj) st.Fprev[jj][f] = (const T*)h->F[jj][f];
        }
        if (int rc = launch_rhs<T>(h, A, st, s)) return rc;
        for (int f = 0; f < 3; ++f) { void* t = Yc[f]; Yc[f] = Yn[f]; Yn[f] = t; }
      }
      return 0;
    };
    auto full_step = [&](double hd, double hnext) -> int {
      if (int rc = stages(hd)) return rc;
      Stage<T> st = empty_stage<T>();
      st.dt = (T)hnext; st.a_new = (T)TSIT5_A[0][0];
      for (int f = 0; f < 3; ++f) {
        st.Yin[f] = (const T*)Yc[f]; st.Fout[f] = (T*)h->F[0][f]; st.Yout[f] = (T*)Yn[f];
      }
      if (int rc = launch_rhs<T>(h, A, st, s)) return rc;
      for (int f = 0; f < 3; ++f) { void* oy = y[f]; y[f] = Yc[f]; Yc[f] = Yn[f]; Yn[f] = oy; }
      return 0;
    };
    long i = 0;
    if (use_graph) {
      StepGraph& G = h->graph;
      const double key[6] = {dt, p->lateral_viscosity, p->bottom_drag, p->wind_amplitude, p->H0, 2.0};
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
      const long pairs = (n_steps - 1) / 2;
      for (long r = 0; r < pairs; ++r) SB_CUDA(cudaGraphLaunch(G.exec, s));
      g_launches.fetch_add(G.nlaunch * (uint64_t)pairs);
      i = 2 * pairs;
    }
    for (; i + 1 < total; ++i)
      if (int rc = full_step(step_dt(i), step_dt(i + 1))) return rc;
    if (int rc = stages(step_dt(total - 1))) return rc;
    for (int f = 0; f < 3; ++f) { void* oy = y[f]; y[f] = Yc[f]; Yc[f] = oy; }
  }
  for (int f = 0; f < 3; ++f)
    if (int rc = unpack_field<T>((const T*)y[f], (T*)ext[f], L, s)) return rc;
  if (use_graph) {
    SB_CUDA(cudaEventRecord(h->graph.ev_out, s));
    SB_CUDA(cudaStreamWaitEvent(caller, h->graph.ev_out, 0));
  }
  return 0;
}

template <typename T>
int swm_rhs_impl(somax_b200_swm_t h, const void* hh, const void* u, const void* v, void* dh,
                 void* du, void* dv, const somax_b200_params* p, int apply_bc, cudaStream_t s) {
  const Layout& L = h->L;
  const void* in[3] = {hh, u, v};
  void* out[3] = {dh, du, dv};
  for (int f = 0; f < 3; ++f)
    if (int rc = pack_field<T>((const T*)in[f], (T*)h->Ya[f], L, s)) return rc;
  SwmArgs<T> A = make_args<T>(h, p, apply_bc);
  Stage<T> st = empty_stage<T>();
  for (int f = 0; f < 3; ++f) { st.Yin[f] = (const T*)h->Ya[f]; st.Fout[f] = (T*)h->F[0][f]; }
  if (int rc = launch_rhs<T>(h, A, st, s)) return rc;
  for (int f = 0; f < 3; ++f)
    if (int rc = unpack_field<T>((const T*)h->F[0][f], (T*)out[f], L, s)) return rc;
  return 0;
}

template <typename T>
int swm_bc_impl(somax_b200_swm_t h, const void* hh, const void* u, const void* v, void* ho,
                void* uo, void* vo, cudaStream_t s) {
  const Layout& L = h->L;
  const void* in[3] = {hh, u, v};
  void* out[3] = {ho, uo, vo};
  for (int f = 0; f < 3; ++f)
    if (int rc = pack_field<T>((const T*)in[f], (T*)h->Ya[f], L, s)) return rc;
  if (int rc = bc_inplace<T>(h, h->Ya, s)) return rc;
  for (int f = 0; f < 3; ++f)
    if (int rc = unpack_field<T>((const T*)h->Ya[f], (T*)out[f], L, s)) return rc;
  return 0;
}

template <typename T>
int swm_project_impl(somax_b200_swm_t h, const void* hh, const void* u, const void* v, void* ho,
                     void* uo, void* vo, cudaStream_t s) {
  const Layout& L = h->L;
  const void* in[3] = {hh, u, v};
  void* out[3] = {ho, uo, vo};
  for (int f = 0; f < 3; ++f)
    if (int rc = pack_field<T>((const T*)in[f], (T*)h->Ya[f], L, s)) return rc;
  const T* pin[3] = {(const T*)h->Ya[0], (const T*)h->Ya[1], (const T*)h->Ya[2]};
  T* pout[3] = {(T*)h->P[0], (T*)h->P[1], (T*)h->P[2]};
  if (int rc = project_state<T>(h, pin, pout, 0, s)) return rc;
  for (int f = 0; f < 3; ++f)
    if (int rc = unpack_field<T>((const T*)h->P[f], (T*)out[f], L, s)) return rc;
  return 0;
}

template <typename T>
int swm_diag_impl(somax_b200_swm_t h, const void* hh, const void* u, const void* v, double* out,
                  cudaStream_t s) {
  const Layout& L = h->L;
  SB_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * L.batch * (3 * L.nl + 1), s));
  dim3 b(256), g(std::min((L.Nx + 255) / 256, 64), std::min(L.Ny, 128), L.batch * L.nl);
  prof_begin("swm_diag_kernel", s);
  swm_diag_kernel<T><<<g, b, 0, s>>>((const T*)hh, (const T*)u, (const T*)v, (const T*)h->f,
                                     h->f1d ? 1 : L.Nx, h->f1d ? 0 : 1, L.nl, L.Ny, L.Nx, h->dx,
                                     h->dy, out);
  SB_LAUNCH_CHECK();
  return 0;
}

}  // namespace

#define SB_DISPATCH(h, fn, ...) \
  ((h)->dtype == SOMAX_B200_F32 ? fn<float>(__VA_ARGS__) : fn<double>(__VA_ARGS__))

extern "C" {

int somax_b200_swm_create(somax_b200_swm_t* out, int dtype, int batch, int nl, int ny, int nx,
                          double dx, double dy, int bc, const double* g_prime,
                          const double* f_field, const double* wind_x, const double* wind_y,
                          unsigned spec_flags) {
  if (!out) return fail(SOMAX_B200_ERR_INVALID, "out is null");
  *out = nullptr;
  if (dtype != SOMAX_B200_F32 && dtype != SOMAX_B200_F64)
    return fail(SOMAX_B200_ERR_INVALID, "dtype must be F32 or F64");
  if (batch < 1 || nl < 1 || nl > SWM_MAX_NL || ny < 3 || nx < 3)
    return fail(SOMAX_B200_ERR_INVALID, "need batch>=1, 1<=nl<=8, ny>=3, nx>=3");
  if (bc != SOMAX_B200_BC_PERIODIC && bc != SOMAX_B200_BC_WALL)
    return fail(SOMAX_B200_ERR_INVALID, "bc must be PERIODIC or WALL");
  if (!g_prime || !f_field || !wind_x || !wind_y)
    return fail(SOMAX_B200_ERR_INVALID, "null coefficient pointer");
  if (int rc = require_device()) return rc;
  auto* h = new somax_b200_swm_s();
  h->dtype = dtype; h->L = make_layout(batch, nl, ny, nx); h->ny = ny; h->nx = nx; h->bc = bc;
  h->spec = spec_flags; h->dx = dx; h->dy = dy;
  for (int k = 0; k < SWM_MAX_NL; ++k) h->gprime[k] = k < nl ? g_prime[k] : 0.0;
  const size_t es = dtype == SOMAX_B200_F32 ? 4 : 8;
  const size_t fb = h->L.count() * es;
  int rc = 0;
  auto up = [&](const double* src, void** dst, bool* one) {
    return dtype == SOMAX_B200_F32 ? upload_coef<float>(src, (float**)dst, h->L.Ny, h->L.Nx, one)
                                   : upload_coef<double>(src, (double**)dst, h->L.Ny, h->L.Nx, one);
  };
  rc = up(f_field, &h->f, &h->f1d);
  if (!rc) rc = up(wind_x, &h->wx, &h->wx1d);
  if (!rc) rc = up(wind_y, &h->wy, &h->wy1d);
  std::vector<void**> bufs;
  for (int f = 0; f < 3; ++f) {
    bufs.push_back(&h->y[f]); bufs.push_back(&h->Ya[f]); bufs.push_back(&h->Yb[f]);
    for (int j = 0; j < 5; ++j) bufs.push_back(&h->F[j][f]);
  }
  for (void** bp : bufs) {
    if (rc) break;
    cudaError_t e = cudaMalloc(bp, fb);
    if (e != cudaSuccess) { rc = fail(SOMAX_B200_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e)); break; }
    cudaMemset(*bp, 0, fb);
    h->bytes += fb;
  }
  if (rc) { somax_b200_swm_destroy(h); return rc; }
  *out = h;
  return 0;
}

int somax_b200_swm_destroy(somax_b200_swm_t h) {
  if (!h) return 0;
  h->graph.destroy();
  cudaFree(h->f); cudaFree(h->wx); cudaFree(h->wy);
  for (int f = 0; f < 3; ++f) {
    cudaFree(h->y[f]); cudaFree(h->Ya[f]); cudaFree(h->Yb[f]); cudaFree(h->P[f]);
    for (int j = 0; j < 5; ++j) cudaFree(h->F[j][f]);
  }
  cudaFree(h->pq); cudaFree(h->ppsi);
  qg_solver_destroy(h->proj);
  delete h;
  return 0;
}

int somax_b200_swm_set_projection(somax_b200_swm_t h, double f0, const double* H, const double* Cl2m,
                                  const double* Cm2l, const double* eigenvalues, const double* lambdas,
                                  int solver) {
  if (!h || !H || !Cl2m || !Cm2l || !eigenvalues || !lambdas) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (h->bc != SOMAX_B200_BC_WALL)
    return fail(SOMAX_B200_ERR_INVALID, "the geostrophic projection needs wall boundary conditions");
  if (h->proj) return fail(SOMAX_B200_ERR_INVALID, "projection already set");
  const Layout& L = h->L;
  const int nl = L.nl;
  if (nl > 4) return fail(SOMAX_B200_ERR_UNSUPPORTED, "projection supports nl <= 4");
  if (int rc = qg_solver_create(&h->proj, h->dtype, L.batch, nl, h->ny, h->nx, h->dx, h->dy, Cl2m, Cm2l, lambdas,
                                solver, 0)) return rc;
  h->pargs.f0 = f0;
  for (int k = 0; k < nl; ++k) {
    h->pargs.H[k] = H[k];
    for (int c = 0; c < nl; ++c) {
      double a = 0;
      for (int m = 0; m < nl; ++m) a += Cm2l[k * nl + m] * eigenvalues[m] * Cl2m[m * nl + c];
      h->pargs.A[k][c] = a;
    }
  }
  const size_t fb = L.count() * (h->dtype == SOMAX_B200_F32 ? 4 : 8);
  void** bufs[] = {&h->pq, &h->ppsi, &h->P[0], &h->P[1], &h->P[2]};
  for (void** bp : bufs) {
    cudaError_t e = cudaMalloc(bp, fb);
    if (e != cudaSuccess) return fail(SOMAX_B200_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cudaMemset(*bp, 0, fb);
    h->bytes += fb;
  }
  h->bytes += qg_solver_bytes(h->proj);
  return 0;
}

size_t somax_b200_swm_device_bytes(somax_b200_swm_t h) { return h ? h->bytes : 0; }

int somax_b200_swm_apply_bc(somax_b200_swm_t h, const void* hh, const void* u, const void* v,
                            void* ho, void* uo, void* vo, void* stream) {
  if (!h || !hh || !u || !v || !ho || !uo || !vo) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  return SB_DISPATCH(h, swm_bc_impl, h, hh, u, v, ho, uo, vo, (cudaStream_t)stream);
}

int somax_b200_swm_project(somax_b200_swm_t h, const void* hh, const void* u, const void* v,
                           void* ho, void* uo, void* vo, void* stream) {
  if (!h || !hh || !u || !v || !ho || !uo || !vo) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (!h->proj) return fail(SOMAX_B200_ERR_INVALID, "no projection on this handle (somax_b200_swm_set_projection)");
  return SB_DISPATCH(h, swm_project_impl, h, hh, u, v, ho, uo, vo, (cudaStream_t)stream);
}

int somax_b200_swm_rhs(somax_b200_swm_t h, const void* hh, const void* u, const void* v, void* dh,
                       void* du, void* dv, const somax_b200_params* p, int apply_bc, void* stream) {
  if (!h || !hh || !u || !v || !dh || !du || !dv || !p) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  return SB_DISPATCH(h, swm_rhs_impl, h, hh, u, v, dh, du, dv, p, apply_bc, (cudaStream_t)stream);
}

int somax_b200_swm_steps(somax_b200_swm_t h, void* hh, void* u, void* v, long n_steps, double dt,
                         double dt_last, const somax_b200_params* p, void* stream) {
  if (!h || !hh || !u || !v || !p) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (n_steps < 0 || !(dt > 0) || dt_last < 0) return fail(SOMAX_B200_ERR_INVALID, "need n_steps>=0, dt>0, dt_last>=0");
  return SB_DISPATCH(h, swm_steps_impl, h, hh, u, v, n_steps, dt, dt_last, p, true, (cudaStream_t)stream);
}

int somax_b200_swm_resume(somax_b200_swm_t h, void* hh, void* u, void* v, long n_steps, double dt,
                          double dt_last, const somax_b200_params* p, void* stream) {
  if (!h || !hh || !u || !v || !p) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  if (n_steps < 0 || !(dt > 0) || dt_last < 0) return fail(SOMAX_B200_ERR_INVALID, "need n_steps>=0, dt>0, dt_last>=0");
  return SB_DISPATCH(h, swm_steps_impl, h, hh, u, v, n_steps, dt, dt_last, p, false, (cudaStream_t)stream);
}

int somax_b200_swm_diag(somax_b200_swm_t h, const void* hh, const void* u, const void* v,
                        double* out, void* stream) {
  if (!h || !hh || !u || !v || !out) return fail(SOMAX_B200_ERR_INVALID, "null argument");
  return SB_DISPATCH(h, swm_diag_impl, h, hh, u, v, out, (cudaStream_t)stream);
}

}  // extern "C"

#include "swm_slab.cuh"
