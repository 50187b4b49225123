// Shallow-water kernels shared by swm.cu (fp32 and the host side) and swm_f64.cu (the fp64
// instantiation of the reference-order kernel, compiled without FMA contraction).
#pragma once
#include "common.cuh"

namespace sb {

constexpr int SWM_MAX_NL = 8;
constexpr int TY = 8;            // output rows per CTA
constexpr int TXG = 32;          // float4 groups per CTA row (=> 128 columns)
constexpr int TW = TXG * 4 + 2;  // tile width incl. halo
constexpr int TWP = TW + 2;      // padded smem row

template <typename T>
struct SwmArgs {
  Layout L;
  int bc;
  unsigned spec;
  int apply_bc;
  // ylo / yhi: row 0 / row Ny-1 of these arrays is the physical ghost row.  A y-slab of the
  // distributed model clears the flag on the side where a neighbouring slab continues: that row is
  // then a halo row holding the neighbour's data - no boundary condition on it, and "interior"
  // for the interior-only operators.
  int ylo, yhi;
  // Periodic basin in slabs: the data of the physical ghost rows after apply_boundary_conditions
  // (row 0 <- global row Ny-2, row Ny-1 <- global row 1) lives on the opposite edge rank; it arrives
  // in the halo inbox and is read from there: lo_src[f] / hi_src[f] = rows [plane][pitch] of field f,
  // or null (one device: the rows are taken from the array itself).
  const T* lo_src[3];
  const T* hi_src[3];
  T dx, dy, dx2, dy2;
  T idx, idy, idx2, idy2, iH0;   // reciprocals (fast kernel)
  const T* f;  int f_cp, f_xs;
  const T* wx; int wx_cp, wx_xs;
  const T* wy; int wy_cp, wy_xs;
  T gprime[SWM_MAX_NL];
  T H0, nu, kappa, tau0;
};

enum { FH = 0, FU = 1, FV = 2 };

// Value of field `kind` at (j,i) after apply_boundary_conditions, read from the raw plane.
// Periodic: enforce_periodic (rows then columns).  Wall: swm/multilayer.py:386-408.
// is row jj inside the region finitevolx's interior-only operators write?
__device__ __forceinline__ bool swm_row_interior(int jj, int Ny, int ylo, int yhi) {
  return jj >= (ylo ? 1 : 0) && jj <= (yhi ? Ny - 2 : Ny - 1);
}

template <typename T>
__device__ __forceinline__ T swm_bc_value(const T* __restrict__ plane, int kind, int bc, int j,
                                          int i, int Ny, int Nx, int pitch, int ylo = 1, int yhi = 1,
                                          const T* __restrict__ lo_row = nullptr,
                                          const T* __restrict__ hi_row = nullptr) {
  const bool lo = ylo && j == 0, hi = yhi && j == Ny - 1;
  if (bc == SOMAX_B200_BC_PERIODIC) {
    int jj = lo ? Ny - 2 : (hi ? 1 : j);
    int ii = (i == 0) ? Nx - 2 : (i == Nx - 1 ? 1 : i);
    if (lo && lo_row) return lo_row[OFF + ii];
    if (hi && hi_row) return hi_row[OFF + ii];
    return plane[(size_t)jj * pitch + OFF + ii];
  }
  int jj = lo ? 1 : (hi ? Ny - 2 : j);
  int ii = (i == 0) ? 1 : (i == Nx - 1 ? Nx - 2 : i);
  if (kind == FU) {
    if (i == 0 || i >= Nx - 2) return T(0);
    return plane[(size_t)jj * pitch + OFF + i];
  }
  if (kind == FV) {
    if (lo || (yhi && j >= Ny - 2)) return T(0);
    return plane[(size_t)j * pitch + OFF + ii];
  }
  return plane[(size_t)jj * pitch + OFF + ii];
}

template <typename T>
__global__ void __launch_bounds__(TXG* TY)
swm_rhs_kernel(SwmArgs<T> A, Stage<T> st) {
  __shared__ T s_h[TY + 2][TWP];
  __shared__ T s_u[TY + 2][TWP];
  __shared__ T s_v[TY + 2][TWP];
  __shared__ T s_p[TY + 2][TWP];

  const Layout& L = A.L;
  const int Ny = L.Ny, Nx = L.Nx, pitch = L.pitch;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TXG + tx;
  const int g0 = blockIdx.x * TXG;           // first float4 group of the tile
  const int c0 = g0 * 4 - OFF;               // column of the tile's first slot
  const int j0 = blockIdx.y * TY;
  const int b = blockIdx.z;
  const int j = j0 + ty;
  const int g = g0 + tx;
  const bool active = (j < Ny) && (g < L.groups());

  // zero the running pressure sum
  for (int e = tid; e < (TY + 2) * TWP; e += TXG * TY) (&s_p[0][0])[e] = T(0);

  for (int k = 0; k < L.nl; ++k) {
    const size_t plane_off = ((size_t)b * L.nl + k) * L.plane();
    const T* ph = st.Yin[FH] + plane_off;
    const T* pu = st.Yin[FU] + plane_off;
    const T* pv = st.Yin[FV] + plane_off;
    __syncthreads();  // previous layer's compute done before the tiles are overwritten
    for (int e = tid; e < (TY + 2) * TW; e += TXG * TY) {
      int r = e / TW, c = e - r * TW;
      int jj = j0 - 1 + r, ii = c0 - 1 + c;
      T vh = 0, vu = 0, vv = 0;
      if (jj >= 0 && jj < Ny && ii >= 0 && ii < Nx) {
        if (A.apply_bc) {
          const size_t ro = ((size_t)b * L.nl + k) * pitch;
          vh = swm_bc_value(ph, FH, A.bc, jj, ii, Ny, Nx, pitch, A.ylo, A.yhi, A.lo_src[FH] ? A.lo_src[FH] + ro : nullptr, A.hi_src[FH] ? A.hi_src[FH] + ro : nullptr);
          vu = swm_bc_value(pu, FU, A.bc, jj, ii, Ny, Nx, pitch, A.ylo, A.yhi, A.lo_src[FU] ? A.lo_src[FU] + ro : nullptr, A.hi_src[FU] ? A.hi_src[FU] + ro : nullptr);
          vv = swm_bc_value(pv, FV, A.bc, jj, ii, Ny, Nx, pitch, A.ylo, A.yhi, A.lo_src[FV] ? A.lo_src[FV] + ro : nullptr, A.hi_src[FV] ? A.hi_src[FV] + ro : nullptr);
        } else {
          size_t o = (size_t)jj * pitch + OFF + ii;
          vh = ph[o]; vu = pu[o]; vv = pv[o];
        }
      }
      s_h[r][c] = vh; s_u[r][c] = vu; s_v[r][c] = vv;
      s_p[r][c] = s_p[r][c] + A.gprime[k] * vh;   // p_k = cumsum_k(g'_k h_k), full grid
    }
    __syncthreads();
    if (!active) continue;

    T out_h[4], out_u[4], out_v[4];
    const int r = ty + 1;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = c0 + tx * 4 + e;
      const int c = tx * 4 + e + 1;
      T dh = 0, du = 0, dv = 0;
      if (i >= 0 && i < Nx) {
        auto H = [&](int dr, int dc) { return s_h[r + dr][c + dc]; };
        auto U = [&](int dr, int dc) { return s_u[r + dr][c + dc]; };
        auto V = [&](int dr, int dc) { return s_v[r + dr][c + dc]; };
        auto P = [&](int dr, int dc) { return s_p[r + dr][c + dc]; };
        auto inI = [&](int dr, int dc) {
          int jj = j + dr, ii = i + dc;
          return swm_row_interior(jj, Ny, A.ylo, A.yhi) && ii >= 1 && ii <= Nx - 2;
        };
        auto Fc = [&](int dr, int dc) {
          return A.f[(size_t)(j + dr) * A.f_cp + (size_t)(i + dc) * A.f_xs];
        };
        const bool interior = inI(0, 0);
        if (interior) {
          // --- potential vorticity at X points (interior-only, zero ring) ---
          auto qX = [&](int dr, int dc) -> T {
            if (!inI(dr, dc)) return T(0);
            T zeta = (V(dr, dc + 1) - V(dr, dc)) / A.dx - (U(dr + 1, dc) - U(dr, dc)) / A.dy;
            T fX = T(0.25) * (Fc(dr, dc) + Fc(dr, dc + 1) + Fc(dr + 1, dc) + Fc(dr + 1, dc + 1));
            T hX = T(0.25) * (H(dr, dc) + H(dr, dc + 1) + H(dr + 1, dc) + H(dr + 1, dc + 1));
            return (zeta + fX) / hX;
          };
          auto vhV = [&](int dr, int dc) -> T {
            return inI(dr, dc) ? (T(0.5) * (H(dr, dc) + H(dr + 1, dc))) * V(dr, dc) : T(0);
          };
          auto uhU = [&](int dr, int dc) -> T {
            return inI(dr, dc) ? (T(0.5) * (H(dr, dc) + H(dr, dc + 1))) * U(dr, dc) : T(0);
          };
          auto keT = [&](int dr, int dc) -> T {
            if (!inI(dr, dc)) return T(0);
            T u2 = T(0.5) * (U(dr, dc) * U(dr, dc) + U(dr, dc - 1) * U(dr, dc - 1));
            T v2 = T(0.5) * (V(dr, dc) * V(dr, dc) + V(dr - 1, dc) * V(dr - 1, dc));
            return T(0.5) * (u2 + v2);
          };
          const T q00 = qX(0, 0);
          const T qU = T(0.5) * (q00 + qX(-1, 0));
          const T qV = T(0.5) * (q00 + qX(0, -1));
          const T vhU = T(0.25) * (vhV(0, 0) + vhV(0, 1) + vhV(-1, 0) + vhV(-1, 1));
          const T uhVv = T(0.25) * (uhU(0, 0) + uhU(1, 0) + uhU(0, -1) + uhU(1, -1));
          const T P00 = keT(0, 0) + P(0, 0);
          const T P01 = keT(0, 1) + P(0, 1);
          const T P10 = keT(1, 0) + P(1, 0);
          du = qU * vhU - (P01 - P00) / A.dx;
          dv = -qV * uhVv - (P10 - P00) / A.dy;
          // --- mass: -div(h u), first-order upwind ---
          auto fe = [&](int dr, int dc) -> T {
            if (!inI(dr, dc)) return T(0);
            T uu = U(dr, dc);
            return uu * (uu > T(0) ? H(dr, dc) : H(dr, dc + 1));
          };
          auto fn = [&](int dr, int dc) -> T {
            if (!inI(dr, dc)) return T(0);
            T vv = V(dr, dc);
            return vv * (vv > T(0) ? H(dr, dc) : H(dr + 1, dc));
          };
          bool wr = true;
          if (A.spec & SOMAX_B200_SPEC_ADVECTION_REGION2)
            wr = (j >= (A.ylo ? 2 : 0) && j <= (A.yhi ? Ny - 3 : Ny - 1) && i >= 2 && i <= Nx - 3);
          if (wr) dh = -((fe(0, 0) - fe(0, -1)) / A.dx + (fn(0, 0) - fn(-1, 0)) / A.dy);
        }
        // --- wind (top layer, FULL grid incl. ring) ---
        if (k == 0) {
          du = du + (A.tau0 * A.wx[(size_t)j * A.wx_cp + (size_t)i * A.wx_xs]) / A.H0;
          dv = dv + (A.tau0 * A.wy[(size_t)j * A.wy_cp + (size_t)i * A.wy_xs]) / A.H0;
        }
        // --- diffusion (interior-only output) ---
        if (interior) {
          T lu, lv;
          if (A.spec & SOMAX_B200_SPEC_DIFFUSION_FLUX) {
            auto fxU = [&](int dr, int dc) -> T {
              return inI(dr, dc) ? A.nu * ((U(dr, dc + 1) - U(dr, dc)) / A.dx) : T(0); };
            auto fyU = [&](int dr, int dc) -> T {
              return inI(dr, dc) ? A.nu * ((U(dr + 1, dc) - U(dr, dc)) / A.dy) : T(0); };
            auto fxV = [&](int dr, int dc) -> T {
              return inI(dr, dc) ? A.nu * ((V(dr, dc + 1) - V(dr, dc)) / A.dx) : T(0); };
            auto fyV = [&](int dr, int dc) -> T {
              return inI(dr, dc) ? A.nu * ((V(dr + 1, dc) - V(dr, dc)) / A.dy) : T(0); };
            lu = (fxU(0, 0) - fxU(0, -1)) / A.dx + (fyU(0, 0) - fyU(-1, 0)) / A.dy;
            lv = (fxV(0, 0) - fxV(0, -1)) / A.dx + (fyV(0, 0) - fyV(-1, 0)) / A.dy;
          } else {
            lu = A.nu * ((U(0, 1) - T(2) * U(0, 0) + U(0, -1)) / A.dx2 +
                         (U(1, 0) - T(2) * U(0, 0) + U(-1, 0)) / A.dy2);
            lv = A.nu * ((V(0, 1) - T(2) * V(0, 0) + V(0, -1)) / A.dx2 +
                         (V(1, 0) - T(2) * V(0, 0) + V(-1, 0)) / A.dy2);
          }
          du = du + lu;
          dv = dv + lv;
        }
        // --- bottom drag (bottom layer, FULL grid incl. ring) ---
        if (k == L.nl - 1) {
          du = du + (-A.kappa * U(0, 0));
          dv = dv + (-A.kappa * V(0, 0));
        }
      }
      out_h[e] = dh; out_u[e] = du; out_v[e] = dv;
    }
    const size_t idx = plane_off + (size_t)j * pitch + (size_t)g * 4;
    Vec4<T> yin;
    Vec4<T> Fh{out_h[0], out_h[1], out_h[2], out_h[3]};
    Vec4<T> Fu{out_u[0], out_u[1], out_u[2], out_u[3]};
    Vec4<T> Fv{out_v[0], out_v[1], out_v[2], out_v[3]};
    const bool need_yin = (st.Yout[0] != nullptr) && (st.y[0] == nullptr);
    yin = need_yin ? ld4(st.Yin[FH] + idx) : Vec4<T>{0, 0, 0, 0};
    rk_epilogue4(st, FH, idx, yin, Fh);
    yin = need_yin ? ld4(st.Yin[FU] + idx) : Vec4<T>{0, 0, 0, 0};
    rk_epilogue4(st, FU, idx, yin, Fu);
    yin = need_yin ? ld4(st.Yin[FV] + idx) : Vec4<T>{0, 0, 0, 0};
    rk_epilogue4(st, FV, idx, yin, Fv);
  }
}

// fp64 launch of swm_rhs_kernel from its own translation unit (swm_f64.cu, nvcc -fmad=false):
// the north-star fp64 tolerance (1e-12 on u, v) needs the reference's operation order - divisions
// by dx, dy, no reciprocal multiplies - AND no fused multiply-adds (numpy / XLA:CPU do not
// contract).  fp64 is the validation pipeline, not the performance path.
int swm_launch_reference_order_f64(const SwmArgs<double>& A, const Stage<double>& st, dim3 grid, dim3 block,
                                   cudaStream_t s);

}  // namespace sb
