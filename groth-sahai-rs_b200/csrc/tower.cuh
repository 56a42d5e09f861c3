// Extension tower Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3-(1+u)), Fp12 = Fp6[w]/(w^2-v)
// for BLS12-381 (same tower and coefficient order as ark-bls12-381, so an fp12 is laid out
// exactly like arkworks' Fp12: c0.c0.c0, c0.c0.c1, c0.c1.c0, ... 12 x 48 B = 576 B).
//
// Replaces: the ark-ff tower arithmetic under E::pairing / E::multi_pairing
// (reference call sites src/data_structures.rs:484-502, src/generator.rs:116).
//
// Execution model on the GPU: one Fp2 operation is the register-resident unit (<= 3 Fp
// Montgomery products, ~1k IMAD.WIDE); everything wider (Fp6/Fp12 values, Miller state)
// lives in per-thread local memory (L1/L2-resident, lane-interleaved => coalesced) and is
// streamed through the noinline Fp2 kernels.  That keeps registers/thread low enough for
// 8-16 warps per SM and the code size bounded.
#pragma once
#include "constants.cuh"
#include "fp.cuh"
#include "modinv.cuh"

namespace gs {

// ------------------------------------------------------------------ Fp helpers
GS_HD GS_INL void fp_set(fp& r, const uint32_t* w) {
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = w[i];
}
GS_HD GS_INL void fp_one(fp& r) {
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = FP_ONE_MONT(i);
}

// a^(p-2) by 4-bit fixed windows (380 squarings + ~110 products); a = 0 -> 0.  Kept as the independent
// cross-check of the safegcd inversion (tests/test_hostsim.py); the product paths call fp_inv below.
inline GS_HD GS_NOINL void fp_inv_fermat(fp& r, const fp& a) {
  fp tab[16];
  fp_one(tab[0]);
  tab[1] = a;
  for (int i = 2; i < 16; i++) fp::mul(tab[i], tab[i - 1], a);
  fp acc;
  fp_one(acc);
  for (int w = 95; w >= 0; w--) {  // p-2 has 381 bits -> 96 nibbles
    if (w != 95) {
      fp::sqr(acc, acc);
      fp::sqr(acc, acc);
      fp::sqr(acc, acc);
      fp::sqr(acc, acc);
    }
    uint32_t nib = (FP_PM2(w >> 3) >> ((w & 7) * 4)) & 15u;
    if (nib) fp::mul(acc, acc, tab[nib]);
  }
  r = acc;
}
// Montgomery inverse by safegcd division steps (modinv.cuh): ~6.6x cheaper than the Fermat chain on B200 and
// mostly integer-ALU work that overlaps with the IMAD-bound kernels around it.  a = 0 -> 0.
GS_HD GS_INL void fp_inv(fp& r, const fp& a) { fp_inv_sg(r, a); }

// ------------------------------------------------------------------ Fp2
// ONE out-of-line copy of the Fp product, operands and result by value (registers across the call): the Fp2 formulas
// below go through it on the device unless GS_FP2_INLINE is defined (three inlined products = 22 KB per Fp2 product,
// which the G2 thread-per-point kernels paid in instruction-cache misses: stall_no_inst 7 % in k_fixed_commit<G2>)
inline GS_HD GS_NOINL fp fp_mul_ool(fp a, fp b) {
  fp r;
  fp::mul(r, a, b);
  return r;
}
#if defined(__CUDA_ARCH__) && !defined(GS_FP2_INLINE)
#define GS_FP2_MUL(r, a, b) (r) = fp_mul_ool((a), (b))
#else
#define GS_FP2_MUL(r, a, b) fp::mul((r), (a), (b))
#endif

struct fp2 {
  fp c0, c1;

  GS_HD static GS_INL void add(fp2& r, const fp2& a, const fp2& b) {
    fp::add(r.c0, a.c0, b.c0);
    fp::add(r.c1, a.c1, b.c1);
  }
  GS_HD static GS_INL void sub(fp2& r, const fp2& a, const fp2& b) {
    fp::sub(r.c0, a.c0, b.c0);
    fp::sub(r.c1, a.c1, b.c1);
  }
  GS_HD static GS_INL void dbl(fp2& r, const fp2& a) {
    fp::add(r.c0, a.c0, a.c0);
    fp::add(r.c1, a.c1, a.c1);
  }
  GS_HD static GS_INL void neg(fp2& r, const fp2& a) {
    fp::neg(r.c0, a.c0);
    fp::neg(r.c1, a.c1);
  }
  GS_HD static GS_INL void conj(fp2& r, const fp2& a) {
    r.c0 = a.c0;
    fp::neg(r.c1, a.c1);
  }
  // r = a * (1 + u)
  GS_HD static GS_INL void mul_xi(fp2& r, const fp2& a) {
    fp t0, t1;
    fp::sub(t0, a.c0, a.c1);
    fp::add(t1, a.c0, a.c1);
    r.c0 = t0;
    r.c1 = t1;
  }
  GS_HD static GS_NOINL void mul(fp2& r, const fp2& a, const fp2& b) {
    fp t0, t1, t2, s0, s1;
    fp::add(s0, a.c0, a.c1);
    fp::add(s1, b.c0, b.c1);
    GS_FP2_MUL(t0, a.c0, b.c0);
    GS_FP2_MUL(t1, a.c1, b.c1);
    GS_FP2_MUL(t2, s0, s1);
    fp::sub(r.c0, t0, t1);
    fp::sub(t2, t2, t0);
    fp::sub(r.c1, t2, t1);
  }
  GS_HD static GS_NOINL void sqr(fp2& r, const fp2& a) {
    fp s, d, m;
    fp::add(s, a.c0, a.c1);
    fp::sub(d, a.c0, a.c1);
    GS_FP2_MUL(m, a.c0, a.c1);
    GS_FP2_MUL(r.c0, s, d);
    fp::add(r.c1, m, m);
  }
  GS_HD static GS_NOINL void mul_fp(fp2& r, const fp2& a, const fp& b) {
    GS_FP2_MUL(r.c0, a.c0, b);
    GS_FP2_MUL(r.c1, a.c1, b);
  }
  GS_HD static GS_NOINL void inv(fp2& r, const fp2& a) {
    fp n, t;
    fp::sqr(n, a.c0);
    fp::sqr(t, a.c1);
    fp::add(n, n, t);
    fp_inv(n, n);
    fp::mul(r.c0, a.c0, n);
    fp::mul(t, a.c1, n);
    fp::neg(r.c1, t);
  }
  GS_HD GS_INL bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  GS_HD GS_INL bool equals(const fp2& o) const { return c0.equals(o.c0) && c1.equals(o.c1); }
  GS_HD GS_INL void set_zero() {
    c0.set_zero();
    c1.set_zero();
  }
  GS_HD GS_INL void set_one() {
    fp_one(c0);
    c1.set_zero();
  }
};

// uniform static interface for the curve templates (F = fp or fp2)
struct FpOps {
  typedef fp T;
  GS_HD static GS_INL void add(fp& r, const fp& a, const fp& b) { fp::add(r, a, b); }
  GS_HD static GS_INL void sub(fp& r, const fp& a, const fp& b) { fp::sub(r, a, b); }
  GS_HD static GS_INL void dbl(fp& r, const fp& a) { fp::add(r, a, a); }
  GS_HD static GS_INL void neg(fp& r, const fp& a) { fp::neg(r, a); }
  // ONE out-of-line copy of the Montgomery product for all G1 curve formulas: doubling + mixed addition are ~13 KB of
  // code instead of ~130 KB, so the loop body of the thread-per-point kernels fits the 32 KB L1.5 instruction cache
  // instead of streaming from L2 (k_vmsm_partial: 13 % "no instruction" stalls, 94.6 -> 80.4 ms once compact).
  // ONE out-of-line copy of the product per kernel (instruction footprint) with operands and result
  // passed BY VALUE: the device ABI then keeps all 36 limbs in registers across the call, where references forced the
  // caller to park the operands in local memory and the callee to load them (36 LDL + 12 STL per product)
  GS_HD static GS_NOINL fp mul_v(fp a, fp b) {
    fp r;
    fp::mul(r, a, b);
    return r;
  }
  GS_HD static GS_INL void mul(fp& r, const fp& a, const fp& b) { r = mul_v(a, b); }
  GS_HD static GS_INL void sqr(fp& r, const fp& a) { mul(r, a, a); }
  GS_HD static GS_INL void inv(fp& r, const fp& a) { fp_inv(r, a); }
  GS_HD static GS_INL void set_one(fp& r) { fp_one(r); }
};
struct Fp2Ops {
  typedef fp2 T;
  GS_HD static GS_INL void add(fp2& r, const fp2& a, const fp2& b) { fp2::add(r, a, b); }
  GS_HD static GS_INL void sub(fp2& r, const fp2& a, const fp2& b) { fp2::sub(r, a, b); }
  GS_HD static GS_INL void dbl(fp2& r, const fp2& a) { fp2::dbl(r, a); }
  GS_HD static GS_INL void neg(fp2& r, const fp2& a) { fp2::neg(r, a); }
  GS_HD static GS_INL void mul(fp2& r, const fp2& a, const fp2& b) { fp2::mul(r, a, b); }
  GS_HD static GS_INL void sqr(fp2& r, const fp2& a) { fp2::sqr(r, a); }
  GS_HD static GS_INL void inv(fp2& r, const fp2& a) { fp2::inv(r, a); }
  GS_HD static GS_INL void set_one(fp2& r) { r.set_one(); }
};

// ------------------------------------------------------------------ Fp6
struct fp6 {
  fp2 c0, c1, c2;

  GS_HD static GS_INL void add(fp6& r, const fp6& a, const fp6& b) {
    fp2::add(r.c0, a.c0, b.c0);
    fp2::add(r.c1, a.c1, b.c1);
    fp2::add(r.c2, a.c2, b.c2);
  }
  GS_HD static GS_INL void sub(fp6& r, const fp6& a, const fp6& b) {
    fp2::sub(r.c0, a.c0, b.c0);
    fp2::sub(r.c1, a.c1, b.c1);
    fp2::sub(r.c2, a.c2, b.c2);
  }
  GS_HD static GS_INL void neg(fp6& r, const fp6& a) {
    fp2::neg(r.c0, a.c0);
    fp2::neg(r.c1, a.c1);
    fp2::neg(r.c2, a.c2);
  }
  // r = a * v
  GS_HD static GS_INL void mul_v(fp6& r, const fp6& a) {
    fp2 t;
    fp2::mul_xi(t, a.c2);
    r.c2 = a.c1;
    r.c1 = a.c0;
    r.c0 = t;
  }
  // Karatsuba, 6 Fp2 products
  GS_HD static GS_NOINL void mul(fp6& r, const fp6& a, const fp6& b) {
    fp2 v0, v1, v2, t0, t1, t2, s;
    fp2::mul(v0, a.c0, b.c0);
    fp2::mul(v1, a.c1, b.c1);
    fp2::mul(v2, a.c2, b.c2);
    // c0 = v0 + xi((a1+a2)(b1+b2) - v1 - v2)
    fp2::add(t0, a.c1, a.c2);
    fp2::add(s, b.c1, b.c2);
    fp2::mul(t0, t0, s);
    fp2::sub(t0, t0, v1);
    fp2::sub(t0, t0, v2);
    fp2::mul_xi(t0, t0);
    fp2::add(t0, t0, v0);
    // c1 = (a0+a1)(b0+b1) - v0 - v1 + xi v2
    fp2::add(t1, a.c0, a.c1);
    fp2::add(s, b.c0, b.c1);
    fp2::mul(t1, t1, s);
    fp2::sub(t1, t1, v0);
    fp2::sub(t1, t1, v1);
    fp2::mul_xi(s, v2);
    fp2::add(t1, t1, s);
    // c2 = (a0+a2)(b0+b2) - v0 - v2 + v1
    fp2::add(t2, a.c0, a.c2);
    fp2::add(s, b.c0, b.c2);
    fp2::mul(t2, t2, s);
    fp2::sub(t2, t2, v0);
    fp2::sub(t2, t2, v2);
    fp2::add(t2, t2, v1);
    r.c0 = t0;
    r.c1 = t1;
    r.c2 = t2;
  }
  // r = a * (b0 + b1 v)      (5 Fp2 products)
  GS_HD static GS_NOINL void mul_by_01(fp6& r, const fp6& a, const fp2& b0, const fp2& b1) {
    fp2 aa, bb, t1, t2, t3, s;
    fp2::mul(aa, a.c0, b0);
    fp2::mul(bb, a.c1, b1);
    fp2::add(s, a.c1, a.c2);
    fp2::mul(t1, s, b1);
    fp2::sub(t1, t1, bb);
    fp2::mul_xi(t1, t1);
    fp2::add(t1, t1, aa);
    fp2::add(s, a.c0, a.c2);
    fp2::mul(t3, s, b0);
    fp2::sub(t3, t3, aa);
    fp2::add(t3, t3, bb);
    fp2::add(t2, b0, b1);
    fp2::add(s, a.c0, a.c1);
    fp2::mul(t2, t2, s);
    fp2::sub(t2, t2, aa);
    fp2::sub(t2, t2, bb);
    r.c0 = t1;
    r.c1 = t2;
    r.c2 = t3;
  }
  // r = a * (b1 v)           (3 Fp2 products)
  GS_HD static GS_NOINL void mul_by_1(fp6& r, const fp6& a, const fp2& b1) {
    fp2 t0, t1, t2;
    fp2::mul(t0, a.c2, b1);
    fp2::mul_xi(t0, t0);
    fp2::mul(t1, a.c0, b1);
    fp2::mul(t2, a.c1, b1);
    r.c0 = t0;
    r.c1 = t1;
    r.c2 = t2;
  }
  GS_HD static GS_NOINL void inv(fp6& r, const fp6& a) {
    fp2 t0, t1, t2, s, d;
    fp2::sqr(t0, a.c0);
    fp2::mul(s, a.c1, a.c2);
    fp2::mul_xi(s, s);
    fp2::sub(t0, t0, s);  // t0 = a0^2 - xi a1 a2
    fp2::sqr(t1, a.c2);
    fp2::mul_xi(t1, t1);
    fp2::mul(s, a.c0, a.c1);
    fp2::sub(t1, t1, s);  // t1 = xi a2^2 - a0 a1
    fp2::sqr(t2, a.c1);
    fp2::mul(s, a.c0, a.c2);
    fp2::sub(t2, t2, s);  // t2 = a1^2 - a0 a2
    fp2::mul(d, a.c2, t1);
    fp2::mul(s, a.c1, t2);
    fp2::add(d, d, s);
    fp2::mul_xi(d, d);
    fp2::mul(s, a.c0, t0);
    fp2::add(d, d, s);
    fp2::inv(d, d);
    fp2::mul(r.c0, t0, d);
    fp2::mul(r.c1, t1, d);
    fp2::mul(r.c2, t2, d);
  }
  GS_HD GS_INL void set_zero() {
    c0.set_zero();
    c1.set_zero();
    c2.set_zero();
  }
};

// ------------------------------------------------------------------ Fp12
struct fp12 {
  fp6 c0, c1;

  GS_HD GS_INL void set_one() {
    c0.set_zero();
    c1.set_zero();
    fp_one(c0.c0.c0);
  }
  GS_HD bool equals(const fp12& o) const {
    const uint32_t* x = (const uint32_t*)this;
    const uint32_t* y = (const uint32_t*)&o;
    uint32_t d = 0;
    for (int i = 0; i < 144; i++) d |= x[i] ^ y[i];
    return d == 0;
  }
  GS_HD static GS_INL void conj(fp12& r, const fp12& a) {
    r.c0 = a.c0;
    fp6::neg(r.c1, a.c1);
  }
  // Karatsuba over Fp6: 3 Fp6 products = 18 Fp2 products
  GS_HD static GS_NOINL void mul(fp12& r, const fp12& a, const fp12& b) {
    fp6 aa, bb, s, t;
    fp6::mul(aa, a.c0, b.c0);
    fp6::mul(bb, a.c1, b.c1);
    fp6::add(s, a.c0, a.c1);
    fp6::add(t, b.c0, b.c1);
    fp6::mul(s, s, t);
    fp6::sub(s, s, aa);
    fp6::sub(r.c1, s, bb);
    fp6::mul_v(bb, bb);
    fp6::add(r.c0, aa, bb);
  }
  // complex squaring: 2 Fp6 products
  GS_HD static GS_NOINL void sqr(fp12& r, const fp12& a) {
    fp6 ab, s, t;
    fp6::mul(ab, a.c0, a.c1);
    fp6::add(s, a.c0, a.c1);
    fp6::mul_v(t, a.c1);
    fp6::add(t, t, a.c0);
    fp6::mul(s, s, t);  // (a0+a1)(a0+v a1) = a0^2 + v a1^2 + (1+v) a0a1
    fp6::sub(s, s, ab);
    fp6::mul_v(t, ab);
    fp6::sub(r.c0, s, t);
    fp6::add(r.c1, ab, ab);
  }
  // r = a * (c0 + c1 v + c4 v w)   -- the M-twist line shape; 13 Fp2 products
  GS_HD static GS_NOINL void mul_by_014(fp12& r, const fp12& a, const fp2& c0, const fp2& c1, const fp2& c4) {
    fp6 aa, bb, s;
    fp2 o;
    fp6::mul_by_01(aa, a.c0, c0, c1);
    fp6::mul_by_1(bb, a.c1, c4);
    fp2::add(o, c1, c4);
    fp6::add(s, a.c0, a.c1);
    fp6::mul_by_01(s, s, c0, o);
    fp6::sub(s, s, aa);
    fp6::sub(r.c1, s, bb);
    fp6::mul_v(bb, bb);
    fp6::add(r.c0, aa, bb);
  }
  GS_HD static GS_NOINL void inv(fp12& r, const fp12& a) {
    fp6 t0, t1;
    fp6::mul(t0, a.c0, a.c0);
    fp6::mul(t1, a.c1, a.c1);
    fp6::mul_v(t1, t1);
    fp6::sub(t0, t0, t1);
    fp6::inv(t0, t0);
    fp6::mul(r.c0, a.c0, t0);
    fp6::mul(t1, a.c1, t0);
    fp6::neg(r.c1, t1);
  }
  // x -> x^(p^K), K = 1 or 2, via the w-power basis a_i w^i  (a_i in Fp2):
  //   (a_i w^i)^(p^K) = conj^K(a_i) * FROB_K[i] * w^i
  template <int K>
  GS_HD static GS_NOINL void frobenius(fp12& r, const fp12& a) {
    const fp2* src[6] = {&a.c0.c0, &a.c1.c0, &a.c0.c1, &a.c1.c1, &a.c0.c2, &a.c1.c2};
    fp2* dst[6] = {&r.c0.c0, &r.c1.c0, &r.c0.c1, &r.c1.c1, &r.c0.c2, &r.c1.c2};
    for (int i = 0; i < 6; i++) {
      fp2 t = *src[i];
      if (K & 1) fp::neg(t.c1, t.c1);
      if (i > 0) {
        fp2 g;
        for (int j = 0; j < 12; j++) {
          g.c0.l[j] = frob_coeff(K, i, 0, j);
          g.c1.l[j] = frob_coeff(K, i, 1, j);
        }
        fp2::mul(t, t, g);
      }
      *dst[i] = t;
    }
  }
  // Granger-Scott squaring for elements of the cyclotomic subgroup (9 Fp2 products-equivalent)
  GS_HD static GS_NOINL void cyclotomic_sqr(fp12& r, const fp12& a) {
    fp2 t0, t1, t2, t3, t4, t5, tmp, s, x;
    // (z0 + z1 y)^2 with z0 = c0.c0, z1 = c1.c1
    sq2(t0, t1, a.c0.c0, a.c1.c1);
    // (z2 + z3 y)^2 with z2 = c1.c0, z3 = c0.c2
    sq2(t2, t3, a.c1.c0, a.c0.c2);
    // (z4 + z5 y)^2 with z4 = c0.c1, z5 = c1.c2
    sq2(t4, t5, a.c0.c1, a.c1.c2);
    // z0 = 3 t0 - 2 z0
    fp2::sub(x, t0, a.c0.c0);
    fp2::dbl(x, x);
    fp2::add(s, x, t0);
    // z1 = 3 t1 + 2 z1
    fp2::add(x, t1, a.c1.c1);
    fp2::dbl(x, x);
    fp2::add(tmp, x, t1);
    r.c0.c0 = s;
    r.c1.c1 = tmp;
    // z2 = 3 xi t5 + 2 z2
    fp2::mul_xi(tmp, t5);
    fp2::add(x, tmp, a.c1.c0);
    fp2::dbl(x, x);
    fp2::add(s, x, tmp);
    // z3 = 3 t4 - 2 z3
    fp2::sub(x, t4, a.c0.c2);
    fp2::dbl(x, x);
    fp2::add(tmp, x, t4);
    r.c1.c0 = s;
    r.c0.c2 = tmp;
    // z4 = 3 t2 - 2 z4
    fp2::sub(x, t2, a.c0.c1);
    fp2::dbl(x, x);
    fp2::add(s, x, t2);
    // z5 = 3 t3 + 2 z5
    fp2::add(x, t3, a.c1.c2);
    fp2::dbl(x, x);
    fp2::add(tmp, x, t3);
    r.c0.c1 = s;
    r.c1.c2 = tmp;
  }

 private:
  // (a + b y)^2 = (a^2 + xi b^2) + (2ab) y   in Fp4 = Fp2[y]/(y^2 - xi)
  GS_HD static GS_INL void sq2(fp2& lo, fp2& hi, const fp2& a, const fp2& b) {
    fp2 ab, s, t;
    fp2::mul(ab, a, b);
    fp2::add(s, a, b);
    fp2::mul_xi(t, b);
    fp2::add(t, t, a);
    fp2::mul(s, s, t);
    fp2::sub(s, s, ab);
    fp2::mul_xi(t, ab);
    fp2::sub(lo, s, t);
    fp2::dbl(hi, ab);
  }
};

}  // namespace gs
