// In-place radix-8/4/2 DIF FFT on a (padded) shared-memory line, and the real-odd split that
// turns a length-n complex FFT into a DST-I of size n-1 (n a power of two).
//
//   X_k = sum_{t=1}^{n-1} x_t sin(pi t k / n),  k = 1..n-1
//
// Method: z = odd extension of x to length 2n (z_t = x_t, z_{2n-t} = -x_t, z_0 = z_n = 0);
// c_t = z_{2t} + i z_{2t+1}; C = FFT_n(c); with A = C_k, B = C_{n-k}:
//   E = (A + conj B)/2, O = (A - conj B)/(2i), Z_k = E + exp(-i pi k/n) O, X_k = -Im(Z_k)/2.
// The shared array viewed as reals IS z (cf[t] = z_t), so loading is two scalar stores per
// input element.  Functions are __host__ __device__ so the index algebra is unit-tested on the
// CPU (tests/test_host_fft.py) without a GPU.
#pragma once
#include <cuda_runtime.h>

namespace sb {

template <typename T> struct C2 { T x, y; };
template <> struct __align__(8) C2<float> { float x, y; };      // one 64-bit / 128-bit access
template <> struct __align__(16) C2<double> { double x, y; };

template <typename T>
__host__ __device__ __forceinline__ C2<T> cadd(C2<T> a, C2<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T>
__host__ __device__ __forceinline__ C2<T> csub(C2<T> a, C2<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T>
__host__ __device__ __forceinline__ C2<T> cmul(C2<T> a, C2<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
__host__ __device__ __forceinline__ C2<T> cscale(C2<T> a, T h) { return {h * a.x, h * a.y}; }
// multiply by -i
template <typename T>
__host__ __device__ __forceinline__ C2<T> mul_mi(C2<T> a) { return {a.y, -a.x}; }

// fp32 on the device: a C2<float> is one 64-bit register pair, so complex add / sub are single
// packed instructions (add.rn.f32x2 -> FADD2).  Same rounding as the scalar form; FADD2 takes one
// issue slot for both lanes, and the transform kernels are issue bound (tools/micro/ffma2_rate.cu).
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long c2_pack(C2<float> a) {
  unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r;
}
__device__ __forceinline__ C2<float> c2_unpack(unsigned long long r) {
  C2<float> a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r)); return a;
}
__device__ __forceinline__ C2<float> cadd(C2<float> a, C2<float> b) {
  unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c2_pack(a)), "l"(c2_pack(b)));
  return c2_unpack(r);
}
__device__ __forceinline__ C2<float> csub(C2<float> a, C2<float> b) {
  unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(c2_pack(a)), "l"(c2_pack(b)));
  return c2_unpack(r);
}
__device__ __forceinline__ C2<float> cscale(C2<float> a, float h) {
  const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(h, h));
  return {r.x, r.y};
}
// complex products as FMUL2 + FFMA2 (two instructions instead of four; ptxas folds the half swaps,
// scalar broadcasts and single-half negations into operand modifiers: .LO_HI, .F32, .NP).  Same
// roundings as the contracted scalar form: re = fma(a.x, b.x, -(a.y b.y)), im = fma(a.x, b.y, a.y b.x).
__device__ __forceinline__ C2<float> cmul(C2<float> a, C2<float> b) {
  const float2 t = __fmul2_rn(make_float2(a.y, a.y), make_float2(b.y, b.x));
  const float2 r = __ffma2_rn(make_float2(a.x, a.x), make_float2(b.x, b.y), make_float2(-t.x, t.y));
  return {r.x, r.y};
}
// a * (c - i sn) with real scalars c, sn
__device__ __forceinline__ C2<float> cmul_cs(C2<float> a, float c, float sn) {
  const float2 t = __fmul2_rn(make_float2(a.y, a.x), make_float2(sn, sn));
  const float2 r = __ffma2_rn(make_float2(a.x, a.y), make_float2(c, c), make_float2(t.x, -t.y));
  return {r.x, r.y};
}
#endif
template <typename T>
__host__ __device__ __forceinline__ C2<T> cmul_cs(C2<T> a, T c, T sn) {
  return {a.x * c + a.y * sn, a.y * c - a.x * sn};
}

// complex index -> padded complex index.  Three-level padding (one slot per 16, per 128 and per
// 1024 elements) keeps both the natural-stride butterfly accesses and the digit-reversed reads
// of the real-odd split spread over the shared-memory banks (simulated: 16 -> 5 wavefronts per
// request for the split at n = 8192, butterflies unchanged at 2.4).
__host__ __device__ __forceinline__ constexpr int fft_pad(int i) { return i + (i >> 4) + (i >> 7) + (i >> 10); }
__host__ __device__ __forceinline__ constexpr int fft_padded_len(int n) { return fft_pad(n) + 1; }

struct FftPlan {
  int n;          // complex length (power of two)
  int lgn;
  int npass;
  int lgr[8];     // log2 of the radix of each pass (3, 2 or 1)
};

inline FftPlan make_fft_plan(int n) {
  FftPlan p;
  p.n = n; p.npass = 0;
  int lg = 0;
  while ((1 << lg) < n) ++lg;
  p.lgn = lg;
  for (int i = 0; i < 8; ++i) p.lgr[i] = 0;
  while (lg >= 3) { p.lgr[p.npass++] = 3; lg -= 3; }
  if (lg == 2) p.lgr[p.npass++] = 2;
  if (lg == 1) p.lgr[p.npass++] = 1;
  return p;
}

// position in the DIF output array of frequency k (mixed-radix digit reversal, shifts only)
__host__ __device__ __forceinline__ int fft_revpos(const FftPlan& p, int k) {
  int pos = 0, sh = p.lgn;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i < p.npass) {
      const int lr = p.lgr[i];
      sh -= lr;
      pos += (k & ((1 << lr) - 1)) << sh;
      k >>= lr;
    }
  }
  return pos;
}

template <typename T>
__host__ __device__ __forceinline__ void fft4(C2<T>& a0, C2<T>& a1, C2<T>& a2, C2<T>& a3) {
  C2<T> t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mi(csub(a1, a3));
  a0 = cadd(t0, t2); a2 = csub(t0, t2); a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

template <typename T>
__host__ __device__ __forceinline__ void fft8(C2<T>* v) {
  // even / odd radix-4
  fft4(v[0], v[2], v[4], v[6]);
  fft4(v[1], v[3], v[5], v[7]);
  const T h = (T)0.70710678118654752440;
  // * exp(-i pi/4), * -i, * exp(-3i pi/4): (v - i v) h and (-i v - v) h, one packed add + one packed multiply
  C2<T> o1 = cadd(v[3], mul_mi(v[3]));
  o1 = cscale(o1, h);
  C2<T> o2 = mul_mi(v[5]);
  C2<T> o3 = csub(mul_mi(v[7]), v[7]);
  o3 = cscale(o3, h);
  C2<T> e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
  v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
  v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
  v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
  v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

// One DIF pass of radix R over sub-transform length Lc for the butterflies t = lt, lt+G, ...
// tw holds exp(-i pi t / n) for t = 0..2n-1 (so w_n^q = tw[2q]).
template <typename T, int R>
__host__ __device__ __forceinline__ void fft_dif_pass(C2<T>* s, int lgn, int lgLc, int lt, int G,
                                                      const C2<T>* __restrict__ tw) {
  constexpr int LGR = R == 8 ? 3 : (R == 4 ? 2 : 1);
  const int lgM = lgLc - LGR;
  const int M = 1 << lgM;
  const int tstride = 2 << (lgn - lgLc);
  const int nb = 1 << (lgn - LGR);
  for (int t = lt; t < nb; t += G) {
    const int block = t >> lgM, pos = t & (M - 1);
    const int base = (block << lgLc) + pos;
    C2<T> v[R];
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = s[fft_pad(base + q * M)];
    if (R == 8) fft8(v);
    else if (R == 4) fft4(v[0], v[1], v[2], v[3]);
    else { C2<T> a = v[0]; v[0] = cadd(a, v[1]); v[1] = csub(a, v[1]); }
    if (M > 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) v[q] = cmul(v[q], tw[(q * pos) * tstride]);
    }
#pragma unroll
    for (int q = 0; q < R; ++q) s[fft_pad(base + q * M)] = v[q];
  }
}

// DST-I post-processing for the pair (k, n-k), 1 <= k <= n/2.
template <typename T>
__host__ __device__ __forceinline__ void dst_split(const C2<T>* s, const FftPlan& p, int k,
                                                   const C2<T>* __restrict__ tw, T& Xk, T& Xnk) {
  const C2<T> A = s[fft_pad(fft_revpos(p, k))];
  const C2<T> B = s[fft_pad(fft_revpos(p, p.n - k))];
  const C2<T> E = {(T)0.5 * (A.x + B.x), (T)0.5 * (A.y - B.y)};
  const C2<T> O = {(T)0.5 * (A.y + B.y), (T)-0.5 * (A.x - B.x)};
  const C2<T> wO = cmul(tw[k], O);
  Xk = (T)-0.5 * (E.y + wO.y);
  Xnk = (T)0.5 * (E.y - wO.y);
}

// ---- compile-time sized variants used by the row kernels (all index algebra constant-folds) ----
template <int LGN> struct FftCT {
  static constexpr int n = 1 << LGN;
  static constexpr int npass = (LGN + 2) / 3;
  __host__ __device__ static constexpr int lgr(int i) {
    return (i < LGN / 3) ? 3 : (LGN % 3);   // radix-8 passes first, then one radix-4 or radix-2
  }
  __host__ __device__ static __forceinline__ int revpos(int k) {
    int pos = 0, sh = LGN;
#pragma unroll
    for (int i = 0; i < npass; ++i) {
      const int lr = lgr(i);
      sh -= lr;
      pos += (k & ((1 << lr) - 1)) << sh;
      k >>= lr;
    }
    return pos;
  }
};

// Compact per-pass twiddle table: for the pass whose butterflies span Lc = 2^LGLC (M = Lc/R) the
// records {w^pos, w^(2 pos), w^(4 pos)}, w = exp(-2 pi i / Lc), pos = 0..M-1, are contiguous, so a
// warp (consecutive pos) reads them coalesced instead of gathering from the length-2n table at
// stride n/Lc.  Every pass before the last is radix 8, so the record offset of a pass is
// 3 * sum of the M of the earlier passes.
__host__ __device__ constexpr int twc_offset(int LGN, int LGLC) {
  int off = 0;
  for (int l = LGN; l > LGLC; l -= 3) off += 3 * (1 << (l - 3));
  return off;
}
__host__ __device__ constexpr int twc_size(int LGN) {
  int off = 0;
  for (int l = LGN; l > 0; l -= (l >= 3 ? 3 : l)) off += 3 * (1 << (l - (l >= 3 ? 3 : l)));
  return off;
}

// DIF pass with compile-time geometry; twiddles w, w^2, w^4 are loaded, the rest are products.
template <typename T, int LGN, int LGLC, int R, int G>
__host__ __device__ __forceinline__ void fft_dif_pass_ct(C2<T>* s, int lt, const C2<T>* __restrict__ twc) {
  constexpr int LGR = R == 8 ? 3 : (R == 4 ? 2 : 1);
  constexpr int lgM = LGLC - LGR;
  constexpr int M = 1 << lgM;
  constexpr int nb = 1 << (LGN - LGR);
  // The butterfly's elements sit at base + q*M.  When, for every pad period 2^s (s = 4, 7, 10),
  // M is a multiple of the period or the whole butterfly span Lc fits inside one period, the
  // padding terms of fft_pad distribute over q and the padded addresses are pad(base) + q*MP
  // with a compile-time MP (true for every pass of the 8192-point transform); otherwise each
  // address is padded individually.
  constexpr int Lc = 1 << LGLC;
  constexpr bool D4 = (M % 16 == 0) || (Lc <= 16);
  constexpr bool D7 = (M % 128 == 0) || (Lc <= 128);
  constexpr bool D10 = (M % 1024 == 0) || (Lc <= 1024);
  constexpr bool DIST = D4 && D7 && D10;
  constexpr int MP = M + (M % 16 == 0 ? M >> 4 : 0) + (M % 128 == 0 ? M >> 7 : 0) + (M % 1024 == 0 ? M >> 10 : 0);
#pragma unroll
  for (int t0 = 0; t0 < nb; t0 += G) {
    const int t = t0 + lt;
    if (nb % G != 0 && t >= nb) break;
    const int block = t >> lgM, pos = t & (M - 1);
    const int base = (block << LGLC) + pos;
    const int pb = fft_pad(base);
    C2<T> v[R];
#pragma unroll
    for (int q = 0; q < R; ++q) v[q] = s[DIST ? pb + q * MP : fft_pad(base + q * M)];
    if (R == 8) fft8(v);
    else if (R == 4) fft4(v[0], v[1], v[2], v[3]);
    else { C2<T> a = v[0]; v[0] = cadd(a, v[1]); v[1] = csub(a, v[1]); }
    if (M > 1) {
      const C2<T>* rec = twc + twc_offset(LGN, LGLC) + 3 * pos;
      const C2<T> w1 = rec[0];
      v[1] = cmul(v[1], w1);
      if (R >= 4) {
        const C2<T> w2 = rec[1];
        const C2<T> w3 = cmul(w1, w2);
        v[2] = cmul(v[2], w2);
        v[3] = cmul(v[3], w3);
        if (R == 8) {
          const C2<T> w4 = rec[2];
          v[4] = cmul(v[4], w4);
          v[5] = cmul(v[5], cmul(w1, w4));
          v[6] = cmul(v[6], cmul(w2, w4));
          v[7] = cmul(v[7], cmul(w3, w4));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) s[DIST ? pb + q * MP : fft_pad(base + q * M)] = v[q];
  }
}

template <typename T, int LGN>
__host__ __device__ __forceinline__ void dst_split_ct(const C2<T>* s, int k,
                                                      const C2<T>* __restrict__ tw, T& Xk, T& Xnk) {
  const C2<T> A = s[fft_pad(FftCT<LGN>::revpos(k))];
  const C2<T> B = s[fft_pad(FftCT<LGN>::revpos((1 << LGN) - k))];
  const C2<T> E = {(T)0.5 * (A.x + B.x), (T)0.5 * (A.y - B.y)};
  const C2<T> O = {(T)0.5 * (A.y + B.y), (T)-0.5 * (A.x - B.x)};
  const C2<T> wO = cmul(tw[k], O);
  Xk = (T)-0.5 * (E.y + wO.y);
  Xnk = (T)0.5 * (E.y - wO.y);
}

// ---- register-resident radix-16/32 butterflies (three-pass transform for n >= 4096) ----
// w_32^m = exp(-2 pi i m / 32) = (FFT_COS32[m], -FFT_SIN32[m]); indices are compile-time after
// unrolling, so the values fold into immediates.
template <typename T>
__host__ __device__ __forceinline__ C2<T> cmul_w32(C2<T> a, int m) {
  constexpr double COS32[32] = {
      1, 0.98078528040323043, 0.92387953251128674, 0.83146961230254524, 0.70710678118654757,
      0.55557023301960229, 0.38268343236508984, 0.19509032201612833, 0, -0.19509032201612819,
      -0.38268343236508973, -0.55557023301960196, -0.70710678118654746, -0.83146961230254535,
      -0.92387953251128674, -0.98078528040323043, -1, -0.98078528040323043, -0.92387953251128685,
      -0.83146961230254546, -0.70710678118654768, -0.55557023301960218, -0.38268343236509034,
      -0.19509032201612866, 0, 0.1950903220161283, 0.38268343236509, 0.55557023301960184,
      0.70710678118654735, 0.83146961230254524, 0.92387953251128652, 0.98078528040323032};
  constexpr double SIN32[32] = {
      0, 0.19509032201612825, 0.38268343236508978, 0.55557023301960218, 0.70710678118654746,
      0.83146961230254524, 0.92387953251128674, 0.98078528040323043, 1, 0.98078528040323043,
      0.92387953251128674, 0.83146961230254546, 0.70710678118654757, 0.55557023301960218,
      0.38268343236508989, 0.19509032201612861, 0, -0.19509032201612836, -0.38268343236508967,
      -0.55557023301960196, -0.70710678118654746, -0.83146961230254524, -0.92387953251128652,
      -0.98078528040323032, -1, -0.98078528040323043, -0.92387953251128663, -0.83146961230254546,
      -0.70710678118654768, -0.55557023301960218, -0.38268343236509039, -0.19509032201612872};
  m &= 31;
  if (m == 0) return a;
  if (m == 8) return mul_mi(a);
  if (m == 16) return {-a.x, -a.y};
  if (m == 24) return {-a.y, a.x};
  return cmul_cs(a, (T)COS32[m], (T)SIN32[m]);
}

// In-register DIF FFT of R = 4 * Rb points (R = 8, 16, 32), natural-order input in v[0..R-1].
// Output frequency k ends up in v[fft_reg_pos<R>(k)].
template <int R>
__host__ __device__ __forceinline__ constexpr int fft_reg_pos(int k) { return (R / 4) * (k & 3) + (k >> 2); }

template <typename T, int R>
__host__ __device__ __forceinline__ void fft_reg(C2<T>* v) {
  constexpr int Rb = R / 4;
  static_assert(R == 8 || R == 16 || R == 32, "radix 8, 16 or 32");
#pragma unroll
  for (int j = 0; j < Rb; ++j) fft4(v[j], v[j + Rb], v[j + 2 * Rb], v[j + 3 * Rb]);
#pragma unroll
  for (int j = 1; j < Rb; ++j)
#pragma unroll
    for (int t = 1; t < 4; ++t) v[j + Rb * t] = cmul_w32(v[j + Rb * t], j * t * (32 / R));
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    if constexpr (Rb == 8) fft8(&v[8 * t]);
    else if constexpr (Rb == 4) fft4(v[4 * t], v[4 * t + 1], v[4 * t + 2], v[4 * t + 3]);
    else { C2<T> a = v[2 * t]; v[2 * t] = cadd(a, v[2 * t + 1]); v[2 * t + 1] = csub(a, v[2 * t + 1]); }
  }
}

// Multiply output frequency q' (q' = 1..R-1, held in v[fft_reg_pos<R>(q')]) by W^{q'} given the
// powers wp[j] = W^(2^j), j < log2 R; the rest are products (at most two per element).
template <typename T, int R>
__host__ __device__ __forceinline__ void fft_reg_twiddle(C2<T>* v, const C2<T>* wp) {
  constexpr int LG = R == 32 ? 5 : (R == 16 ? 4 : 3);
  C2<T> lo[8];
  lo[1] = wp[0]; lo[2] = wp[1]; lo[3] = cmul(lo[1], lo[2]); lo[4] = wp[2];
  lo[5] = cmul(lo[1], lo[4]); lo[6] = cmul(lo[2], lo[4]); lo[7] = cmul(lo[3], lo[4]);
#pragma unroll
  for (int q = 1; q < 8; ++q) v[fft_reg_pos<R>(q)] = cmul(v[fft_reg_pos<R>(q)], lo[q]);
  if constexpr (LG >= 4) {
#pragma unroll
    for (int h = 1; h < R / 8; ++h) {
      C2<T> wh = (h == 1) ? wp[3] : (h == 2 ? wp[LG >= 5 ? 4 : 3] : cmul(wp[3], wp[LG >= 5 ? 4 : 3]));
      v[fft_reg_pos<R>(8 * h)] = cmul(v[fft_reg_pos<R>(8 * h)], wh);
#pragma unroll
      for (int l = 1; l < 8; ++l)
        v[fft_reg_pos<R>(8 * h + l)] = cmul(v[fft_reg_pos<R>(8 * h + l)], cmul(wh, lo[l]));
    }
  }
}

}  // namespace sb
