// Miller accumulation, v2: the Fp12 accumulator lives in SHARED memory in the w-power basis
//     f = sum_{i<6} a_i w^i ,  a_i in Fp2   (a_0=c0.c0, a_1=c1.c0, a_2=c0.c1, a_3=c1.c1, a_4=c0.c2, a_5=c1.c2)
// laid out word-major / thread-minor (word j of coefficient i of thread t at sm[(i*24 + j)*NT + t]),
// which is bank-conflict free, and every update is done IN PLACE one output coefficient at a time with
// register-only temporaries:
//   * line multiplication  f <- f * (c0 + c1 w^2 + c4 w^3)   (M-twist line, = mul_by_014):
//       b_k = c0 a_k + c1 a_{k-2} + c4 a_{k-3}   (x xi on wrap-around), 18 Fp2 products, 3 saved inputs
//   * squaring  f <- f^2 : symmetric schoolbook, 6 Fp2 squarings + 15 Fp2 products
// No local-memory temporaries remain (v1 spilled ~2.7 KB/thread and moved 150 GB of DRAM traffic per
// launch, profiles/r01_ncu_full_summary.csv).  The price is 18 instead of 13 Fp2 products per line.
#pragma once
#include "pairing.cuh"

namespace gs {

// strided Fp2 load/store (stride in 32-bit words between consecutive limbs)
GS_HD GS_INL void ld_fp2(fp2& r, const uint32_t* p, int stride) {
#pragma unroll
  for (int j = 0; j < 12; j++) {
    r.c0.l[j] = p[j * stride];
    r.c1.l[j] = p[(12 + j) * stride];
  }
}
GS_HD GS_INL void st_fp2(uint32_t* p, int stride, const fp2& a) {
#pragma unroll
  for (int j = 0; j < 12; j++) {
    p[j * stride] = a.c0.l[j];
    p[(12 + j) * stride] = a.c1.l[j];
  }
}
GS_HD GS_INL void fp2_mul_inl(fp2& r, const fp2& a, const fp2& b) {
  fp t0, t1, t2, s0, s1;
  fp::add(s0, a.c0, a.c1);
  fp::add(s1, b.c0, b.c1);
  fp::mul(t0, a.c0, b.c0);
  fp::mul(t1, a.c1, b.c1);
  fp::mul(t2, s0, s1);
  fp::sub(r.c0, t0, t1);
  fp::sub(t2, t2, t0);
  fp::sub(r.c1, t2, t1);
}
GS_HD GS_INL void fp2_sqr_inl(fp2& r, const fp2& a) {
  fp s, d, m;
  fp::add(s, a.c0, a.c1);
  fp::sub(d, a.c0, a.c1);
  fp::mul(m, a.c0, a.c1);
  fp::mul(r.c0, s, d);
  fp::add(r.c1, m, m);
}

#define GS_COEF(base, i, stride) ((base) + (size_t)(i) * 24 * (stride))

// f <- f * (c0 + c1 w^2 + c4 w^3); f and lc (c0,c1,c4 = 3 Fp2) are strided arrays
GS_HD GS_INL void f12w_mul_line(uint32_t* f, const uint32_t* lc, int stride) {
  fp2 s3, s4, s5, a, c, acc, t;
  ld_fp2(s3, GS_COEF(f, 3, stride), stride);
  ld_fp2(s4, GS_COEF(f, 4, stride), stride);
  ld_fp2(s5, GS_COEF(f, 5, stride), stride);
  // b5 = c0 a5 + c1 a3 + c4 a2
  ld_fp2(c, GS_COEF(lc, 0, stride), stride);
  fp2_mul_inl(acc, s5, c);
  ld_fp2(c, GS_COEF(lc, 1, stride), stride);
  fp2_mul_inl(t, s3, c);
  fp2::add(acc, acc, t);
  ld_fp2(c, GS_COEF(lc, 2, stride), stride);
  ld_fp2(a, GS_COEF(f, 2, stride), stride);
  fp2_mul_inl(t, a, c);
  fp2::add(acc, acc, t);
  st_fp2(GS_COEF(f, 5, stride), stride, acc);
  // b4 = c0 a4 + c1 a2 + c4 a1
  ld_fp2(c, GS_COEF(lc, 1, stride), stride);
  fp2_mul_inl(acc, a, c);  // a = a2
  ld_fp2(c, GS_COEF(lc, 0, stride), stride);
  fp2_mul_inl(t, s4, c);
  fp2::add(acc, acc, t);
  ld_fp2(c, GS_COEF(lc, 2, stride), stride);
  ld_fp2(a, GS_COEF(f, 1, stride), stride);
  fp2_mul_inl(t, a, c);
  fp2::add(acc, acc, t);
  st_fp2(GS_COEF(f, 4, stride), stride, acc);
  // b3 = c0 a3 + c1 a1 + c4 a0
  ld_fp2(c, GS_COEF(lc, 1, stride), stride);
  fp2_mul_inl(acc, a, c);  // a = a1
  ld_fp2(c, GS_COEF(lc, 0, stride), stride);
  fp2_mul_inl(t, s3, c);
  fp2::add(acc, acc, t);
  ld_fp2(c, GS_COEF(lc, 2, stride), stride);
  ld_fp2(a, GS_COEF(f, 0, stride), stride);
  fp2_mul_inl(t, a, c);
  fp2::add(acc, acc, t);
  st_fp2(GS_COEF(f, 3, stride), stride, acc);
  // b2 = c0 a2 + c1 a0 + xi c4 a5
  fp2_mul_inl(acc, s5, c);  // c = c4
  fp2::mul_xi(acc, acc);
  ld_fp2(c, GS_COEF(lc, 1, stride), stride);
  fp2_mul_inl(t, a, c);  // a = a0
  fp2::add(acc, acc, t);
  ld_fp2(c, GS_COEF(lc, 0, stride), stride);
  ld_fp2(a, GS_COEF(f, 2, stride), stride);
  fp2_mul_inl(t, a, c);
  fp2::add(acc, acc, t);
  st_fp2(GS_COEF(f, 2, stride), stride, acc);
  // b1 = c0 a1 + xi (c1 a5 + c4 a4)
  ld_fp2(c, GS_COEF(lc, 1, stride), stride);
  fp2_mul_inl(acc, s5, c);
  ld_fp2(c, GS_COEF(lc, 2, stride), stride);
  fp2_mul_inl(t, s4, c);
  fp2::add(acc, acc, t);
  fp2::mul_xi(acc, acc);
  ld_fp2(c, GS_COEF(lc, 0, stride), stride);
  ld_fp2(a, GS_COEF(f, 1, stride), stride);
  fp2_mul_inl(t, a, c);
  fp2::add(acc, acc, t);
  st_fp2(GS_COEF(f, 1, stride), stride, acc);
  // b0 = c0 a0 + xi (c1 a4 + c4 a3)
  ld_fp2(c, GS_COEF(lc, 1, stride), stride);
  fp2_mul_inl(acc, s4, c);
  ld_fp2(c, GS_COEF(lc, 2, stride), stride);
  fp2_mul_inl(t, s3, c);
  fp2::add(acc, acc, t);
  fp2::mul_xi(acc, acc);
  ld_fp2(c, GS_COEF(lc, 0, stride), stride);
  ld_fp2(a, GS_COEF(f, 0, stride), stride);
  fp2_mul_inl(t, a, c);
  fp2::add(acc, acc, t);
  st_fp2(GS_COEF(f, 0, stride), stride, acc);
}

// acc += 2 * a_i * a_j (i != j), both read from f
GS_HD GS_INL void f12w_cross(fp2& acc, const uint32_t* f, int i, int j, int stride, bool first) {
  fp2 x, y, t;
  ld_fp2(x, GS_COEF(f, i, stride), stride);
  ld_fp2(y, GS_COEF(f, j, stride), stride);
  fp2_mul_inl(t, x, y);
  fp2::dbl(t, t);
  if (first)
    acc = t;
  else
    fp2::add(acc, acc, t);
}
GS_HD GS_INL void f12w_square_term(fp2& acc, const uint32_t* f, int i, int stride, bool first) {
  fp2 x, t;
  ld_fp2(x, GS_COEF(f, i, stride), stride);
  fp2_sqr_inl(t, x);
  if (first)
    acc = t;
  else
    fp2::add(acc, acc, t);
}

// f <- f^2.  b0..b2 are parked in `tmp` (3 Fp2, strided like f), b3..b5 in registers, then all written back.
GS_HD GS_INL void f12w_sqr(uint32_t* f, uint32_t* tmp, int stride) {
  fp2 acc, hi, b3, b4, b5;
  // b0 = a0^2 + xi (2 a1 a5 + 2 a2 a4 + a3^2)
  f12w_cross(hi, f, 1, 5, stride, true);
  f12w_cross(hi, f, 2, 4, stride, false);
  f12w_square_term(hi, f, 3, stride, false);
  fp2::mul_xi(hi, hi);
  f12w_square_term(acc, f, 0, stride, true);
  fp2::add(acc, acc, hi);
  st_fp2(GS_COEF(tmp, 0, stride), stride, acc);
  // b1 = 2 a0 a1 + xi (2 a2 a5 + 2 a3 a4)
  f12w_cross(hi, f, 2, 5, stride, true);
  f12w_cross(hi, f, 3, 4, stride, false);
  fp2::mul_xi(hi, hi);
  f12w_cross(acc, f, 0, 1, stride, true);
  fp2::add(acc, acc, hi);
  st_fp2(GS_COEF(tmp, 1, stride), stride, acc);
  // b2 = 2 a0 a2 + a1^2 + xi (2 a3 a5 + a4^2)
  f12w_cross(hi, f, 3, 5, stride, true);
  f12w_square_term(hi, f, 4, stride, false);
  fp2::mul_xi(hi, hi);
  f12w_cross(acc, f, 0, 2, stride, true);
  f12w_square_term(acc, f, 1, stride, false);
  fp2::add(acc, acc, hi);
  st_fp2(GS_COEF(tmp, 2, stride), stride, acc);
  // b3 = 2 a0 a3 + 2 a1 a2 + xi (2 a4 a5)
  f12w_cross(hi, f, 4, 5, stride, true);
  fp2::mul_xi(hi, hi);
  f12w_cross(b3, f, 0, 3, stride, true);
  f12w_cross(b3, f, 1, 2, stride, false);
  fp2::add(b3, b3, hi);
  // b4 = 2 a0 a4 + 2 a1 a3 + a2^2 + xi a5^2
  f12w_square_term(hi, f, 5, stride, true);
  fp2::mul_xi(hi, hi);
  f12w_cross(b4, f, 0, 4, stride, true);
  f12w_cross(b4, f, 1, 3, stride, false);
  f12w_square_term(b4, f, 2, stride, false);
  fp2::add(b4, b4, hi);
  // b5 = 2 a0 a5 + 2 a1 a4 + 2 a2 a3
  f12w_cross(b5, f, 0, 5, stride, true);
  f12w_cross(b5, f, 1, 4, stride, false);
  f12w_cross(b5, f, 2, 3, stride, false);
  st_fp2(GS_COEF(f, 3, stride), stride, b3);
  st_fp2(GS_COEF(f, 4, stride), stride, b4);
  st_fp2(GS_COEF(f, 5, stride), stride, b5);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    ld_fp2(acc, GS_COEF(tmp, k, stride), stride);
    st_fp2(GS_COEF(f, k, stride), stride, acc);
  }
}

GS_HD GS_INL void f12w_set_one(uint32_t* f, int stride) {
  for (int w = 0; w < 144; w++) f[(size_t)w * stride] = 0;
  for (int j = 0; j < 12; j++) f[(size_t)j * stride] = FP_ONE_MONT(j);
}
// conjugate (negate the odd w-powers) and write out in tower order
GS_HD GS_INL void f12w_store_conj(fp12& out, const uint32_t* f, int stride) {
  fp2 t;
  ld_fp2(t, GS_COEF(f, 0, stride), stride);
  out.c0.c0 = t;
  ld_fp2(t, GS_COEF(f, 2, stride), stride);
  out.c0.c1 = t;
  ld_fp2(t, GS_COEF(f, 4, stride), stride);
  out.c0.c2 = t;
  ld_fp2(t, GS_COEF(f, 1, stride), stride);
  fp2::neg(out.c1.c0, t);
  ld_fp2(t, GS_COEF(f, 3, stride), stride);
  fp2::neg(out.c1.c1, t);
  ld_fp2(t, GS_COEF(f, 5, stride), stride);
  fp2::neg(out.c1.c2, t);
}

}  // namespace gs
