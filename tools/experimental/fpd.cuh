// Fp sums of products on the FP64 pipe (DFMA) -- EXPERIMENTAL building block, not yet used by the product kernels.
//
// Why: on B200 the FP64 pipe issues 65.7 DFMA/clk/SM and runs fully concurrently with the integer pipe whose
// IMAD.WIDE (26-32/clk/SM) bounds every hot kernel today (profiles/r01c_microbench.txt: DFMA on half of the warps
// next to IMAD.WIDE on the other half costs the IMAD half's time alone).  A Montgomery sum of products that lives on
// the FP64 pipe can therefore run NEXT TO the integer one (different warps of the same block) instead of after it.
//
// Representation: x = sum_{i<8} x_i 2^(48 i), limbs as exact integers in [0, 2^48) held in doubles; 8 x 48 = 384, so
// the Montgomery radix is the same R = 2^384 as the 12 x 32-bit integer code and results are interchangeable.
// Exact 48 x 48 -> 96-bit products from two FMAs in round-toward-zero (Emmart, Zheng, Weems, ARITH 2018):
//     hi = fma_rz(a, b, 2^100)                =  2^100 + floor(ab / 2^48) 2^48      (ulp of that binade is 2^48)
//     lo = fma_rz(a, b, (2^100 + 2^52) - hi)  =  2^52 + (ab mod 2^48)               (exact)
// The BIT PATTERNS of hi / lo are the integers hi48 + bits(2^100) / lo48 + bits(2^52); column sums are accumulated
// as 64-bit integers (IADD3 pairs on the ALU pipe) with the constant biases pre-subtracted.  Both biases are
// multiples of 2^48, so the low 48 bits of a column are right at any time.
// CIOS with 48-bit words: row j adds a[t] * b[t]_j for all terms t, then q_j = col_j * (-p^-1) mod 2^48 and q_j * p;
// per Fp product 64 + 8 + 64/NT-amortised products of 3 FP64 ops each.
//
// tests/hostsim runs the same code on the CPU with fesetround(FE_TOWARDZERO) and compares with fp::mulsum.
#pragma once
#include <math.h>

#include "fp.cuh"

namespace gs {

#if defined(__CUDA_ARCH__)
#define GS_FMA_RZ(a, b, c) __fma_rz((a), (b), (c))
#else
#define GS_FMA_RZ(a, b, c) fma((a), (b), (c))  // host: the caller has set FE_TOWARDZERO
#endif

GS_HD GS_INL double fpd_bits_to_double(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)v);
#else
  double d;
  memcpy(&d, &v, 8);
  return d;
#endif
}
GS_HD GS_INL uint64_t fpd_double_to_bits(double d) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(d);
#else
  uint64_t v;
  memcpy(&v, &d, 8);
  return v;
#endif
}
// exact integer v < 2^52 -> double
GS_HD GS_INL double fpd_u52_to_double(uint64_t v) { return fpd_bits_to_double(0x4330000000000000ull | v) - 4503599627370496.0; }

constexpr double FPD_C1 = 1267650600228229401496703205376.0;                      // 2^100
constexpr double FPD_C2 = 1267650600228229401496703205376.0 + 4503599627370496.0;  // 2^100 + 2^52
constexpr uint64_t FPD_BH = 0x4630000000000000ull;  // bits(2^100)
constexpr uint64_t FPD_BL = 0x4330000000000000ull;  // bits(2^52)
constexpr uint64_t FPD_MASK48 = 0xFFFFFFFFFFFFull;
GS_HD constexpr double fpd_mod48(int i) {  // p in 48-bit limbs
  constexpr double t[8] = {281474976688811.0, 194974335351294.0, 270634993844222.0, 113459389855408.0,
                           83034393350847.0,  73992301405303.0,  253550359455670.0, 28591897852287.0};
  return t[i];
}
constexpr double FPD_NP48 = 281462091612157.0;  // -p^-1 mod 2^48

// 12 x 32-bit limbs -> 8 x 48-bit limbs as doubles
GS_HD GS_INL void fpd_from_fp(double (&d)[8], const fp& a) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint64_t w0 = a.l[3 * j], w1 = a.l[3 * j + 1], w2 = a.l[3 * j + 2];
    d[2 * j] = fpd_u52_to_double(w0 | ((w1 & 0xFFFFull) << 32));
    d[2 * j + 1] = fpd_u52_to_double((w1 >> 16) | (w2 << 16));
  }
}

// number of (i, j) in [0,8)^2 with i + j = k
GS_HD constexpr int fpd_diag(int k) { return k < 0 || k > 14 ? 0 : (k < 8 ? k + 1 : 15 - k); }
// pre-subtracted bias of column k for NT terms: (NT + 1) rows of products land a "lo" pattern on the diagonal
// i + j = k and a "hi" pattern on the diagonal i + j = k - 1
GS_HD constexpr uint64_t fpd_col_bias(int k, int NT) {
  return (uint64_t)0 - ((uint64_t)((NT + 1) * fpd_diag(k)) * FPD_BL + (uint64_t)((NT + 1) * fpd_diag(k - 1)) * FPD_BH);
}

// col[i + j] += lo48(a b) (+ bias), col[i + j + 1] += hi48(a b) (+ bias)
GS_HD GS_INL void fpd_mac(uint64_t& c_lo, uint64_t& c_hi, double a, double b) {
  const double hi = GS_FMA_RZ(a, b, FPD_C1);
  const double lo = GS_FMA_RZ(a, b, FPD_C2 - hi);
  c_lo += fpd_double_to_bits(lo);
  c_hi += fpd_double_to_bits(hi);
}

// r = (sum_t a[t] * b[t]) / R mod p, canonical; inputs canonical (or within the `units` bound of fp::mulsum)
template <int NT>
GS_HD GS_INL void mulsum_dfma(fp& r, const fp (&a)[NT], const fp (&b)[NT]) {
  double A[NT][8], B[NT][8];
#pragma unroll
  for (int t = 0; t < NT; t++) {
    fpd_from_fp(A[t], a[t]);
    fpd_from_fp(B[t], b[t]);
  }
  uint64_t col[17];
#pragma unroll
  for (int k = 0; k < 17; k++) col[k] = fpd_col_bias(k, NT);
#pragma unroll
  for (int j = 0; j < 8; j++) {
#pragma unroll
    for (int t = 0; t < NT; t++) {
#pragma unroll
      for (int i = 0; i < 8; i++) fpd_mac(col[i + j], col[i + j + 1], A[t][i], B[t][j]);
    }
    // column j is complete up to q_j * p_0: its low 48 bits are final
    const double tl = fpd_u52_to_double(col[j] & FPD_MASK48);
    const double qh = GS_FMA_RZ(tl, FPD_NP48, FPD_C1);
    const double q = GS_FMA_RZ(tl, FPD_NP48, FPD_C1 - qh);  // (tl * N') mod 2^48, exact
#pragma unroll
    for (int i = 0; i < 8; i++) fpd_mac(col[i + j], col[i + j + 1], q, fpd_mod48(i));
    col[j + 1] += (uint64_t)((int64_t)col[j] >> 48);  // low 48 bits are zero now; biases are all in
  }
  // columns 8..15 hold the result (< 2p): carry-normalise to 48-bit limbs, repack to 32-bit limbs
  uint64_t limb[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    limb[k] = col[8 + k] & FPD_MASK48;
    col[9 + k] += (uint64_t)((int64_t)col[8 + k] >> 48);
  }
  uint32_t t[12];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint64_t e = limb[2 * j], o = limb[2 * j + 1];
    t[3 * j] = (uint32_t)e;
    t[3 * j + 1] = (uint32_t)(e >> 32) | ((uint32_t)o << 16);
    t[3 * j + 2] = (uint32_t)(o >> 16);
  }
  fp::final_sub(r, t);
}

// ------------------------------------------------------------------ rolled form (small code footprint)
// The unrolled form above is ~2,300 instructions; next to the ~1,000 instructions of fp::mulsum on other warps it
// thrashes the instruction cache (measured: 15 % "no instruction" stalls in the mixed microbenchmark).  Here the 8
// CIOS rows are a real loop over a 9-column sliding window, ~300 instructions per row.  `bw(t, w)` returns 32-bit
// word w of operand b[t] (a shared-memory or local-memory load: the row index is a run-time value).
// Bias bookkeeping is per row: a row lands (NT+1) "lo" patterns on window columns 0..7 and (NT+1) "hi" patterns on
// columns 1..8, so those constants are subtracted when the row starts.
template <int NT, class BW>
GS_HD GS_INL void mulsum_dfma_rolled(fp& r, const fp (&a)[NT], BW&& bw) {
  double A[NT][8];
#pragma unroll
  for (int t = 0; t < NT; t++) fpd_from_fp(A[t], a[t]);
  uint64_t c[9];
#pragma unroll
  for (int k = 0; k < 9; k++) c[k] = 0;
  constexpr uint64_t KL = (uint64_t)(NT + 1) * FPD_BL, KH = (uint64_t)(NT + 1) * FPD_BH;
#pragma unroll 1
  for (int jj = 0; jj < 4; jj++) {  // two 48-bit rows per iteration = three 32-bit words of every b[t]
#pragma unroll
    for (int h = 0; h < 2; h++) {
      c[0] -= KL;
#pragma unroll
      for (int k = 1; k < 8; k++) c[k] -= KL + KH;
      c[8] -= KH;
#pragma unroll
      for (int t = 0; t < NT; t++) {
        const uint64_t w0 = bw(t, 3 * jj), w1 = bw(t, 3 * jj + 1), w2 = bw(t, 3 * jj + 2);
        const double bj = h == 0 ? fpd_u52_to_double(w0 | ((w1 & 0xFFFFull) << 32)) : fpd_u52_to_double((w1 >> 16) | (w2 << 16));
#pragma unroll
        for (int i = 0; i < 8; i++) fpd_mac(c[i], c[i + 1], A[t][i], bj);
      }
      const double tl = fpd_u52_to_double(c[0] & FPD_MASK48);
      const double qh = GS_FMA_RZ(tl, FPD_NP48, FPD_C1);
      const double q = GS_FMA_RZ(tl, FPD_NP48, FPD_C1 - qh);
#pragma unroll
      for (int i = 0; i < 8; i++) fpd_mac(c[i], c[i + 1], q, fpd_mod48(i));
      const uint64_t carry = (uint64_t)((int64_t)c[0] >> 48);
#pragma unroll
      for (int k = 0; k < 8; k++) c[k] = c[k + 1];
      c[0] += carry;
      c[8] = 0;
    }
  }
  uint64_t limb[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    limb[k] = c[k] & FPD_MASK48;
    if (k < 7) c[k + 1] += (uint64_t)((int64_t)c[k] >> 48);
  }
  uint32_t t[12];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const uint64_t e = limb[2 * j], o = limb[2 * j + 1];
    t[3 * j] = (uint32_t)e;
    t[3 * j + 1] = (uint32_t)(e >> 32) | ((uint32_t)o << 16);
    t[3 * j + 2] = (uint32_t)(o >> 16);
  }
  fp::final_sub(r, t);
}

}  // namespace gs
