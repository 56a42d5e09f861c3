// Optimal-ate pairing on BLS12-381, split the B200 way:
//
//   (1) line walk      a thread walks E G2 points in AFFINE coordinates over the 63+5 steps of
//                      |x| = 0xd201000000010000 (g2_affine_step below: one shared safegcd inversion per
//                      step), evaluates every line at the slot's two G1 coordinates and writes the
//                      unit-w^3 tiles that the Miller kernel consumes (pairing.cu k_g2_prepare4).
//   (2) Miller         6 warps = 32 GT accumulators, one warp per w-power coefficient (coop12.cuh):
//                      f <- f^2 * prod_k line_k(P_k), tiles streamed from HBM with cp.async.
//   (3) final exp      easy part + the arkworks hard part (x-1)^2 (x+p)(x^2+p^2-1) + 3 as an op
//                      program over shared-memory buffers (finalexp.cu).
// The homogeneous-projective steps and the thread-per-accumulator tower code below are the first
// generation, kept because tests/hostsim uses them as an independent cross-check of the cooperative path.
//
// Replaces: ark-ec `Bls12::multi_miller_loop` + `final_exponentiation` as reached from
// ComT::pairing / ComT::pairing_sum (src/data_structures.rs:484-502) and generator.rs:116.
// Any Miller function that differs by proper-subfield factors gives identical final bits;
// the final exponent is arkworks' (3x the textbook one), see SURVEY.md §8c.
#pragma once
#include "curve.cuh"

namespace gs {

constexpr int GS_NUM_LINES = 68;  // 63 doublings + 5 additions
constexpr uint64_t GS_X_ABS = 0xd201000000010000ull;

struct line_coeffs {  // ell = c0 + (c1 * xP) v + (c2 * yP) v w      (M-twist => mul_by_014)
  fp2 c0, c1, c2;
};

struct g2_proj {  // homogeneous projective
  fp2 x, y, z;
};

GS_HD GS_INL void fp2_mul_b_twist(fp2& r, const fp2& a) {
  // * 4(1+u):  4 * (a0 - a1, a0 + a1)
  fp2 t;
  fp2::mul_xi(t, a);
  fp2::dbl(t, t);
  fp2::dbl(r, t);
}
GS_HD GS_INL void fp_half(fp& r, const fp& a) {
  // a/2 mod p = (a + (a odd ? p : 0)) >> 1   (the Montgomery representative halves like the value)
  uint32_t mask = 0u - (a.l[0] & 1u);
  uint32_t t[12];
  t[0] = add_cc(a.l[0], FpParams::mod(0) & mask);
#pragma unroll
  for (int i = 1; i < 11; i++) t[i] = addc_cc(a.l[i], FpParams::mod(i) & mask);
  t[11] = addc(a.l[11], FpParams::mod(11) & mask);  // a + p < 2^382: no carry out
#pragma unroll
  for (int i = 0; i < 11; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
  r.l[11] = t[11] >> 1;
}
GS_HD GS_INL void fp2_half(fp2& r, const fp2& a) {
  fp_half(r.c0, a.c0);
  fp_half(r.c1, a.c1);
}

// Costello-Lange-Naehrig doubling step in homogeneous projective coordinates.
inline GS_HD GS_NOINL void g2_double_step(g2_proj& t, line_coeffs& l) {
  fp2 a, b, c, e, f, g, h, i, j, e2, s;
  fp2::mul(a, t.x, t.y);
  fp2_half(a, a);
  fp2::sqr(b, t.y);
  fp2::sqr(c, t.z);
  fp2::dbl(s, c);
  fp2::add(s, s, c);
  fp2_mul_b_twist(e, s);  // e = b' * 3c
  fp2::dbl(f, e);
  fp2::add(f, f, e);  // f = 3e
  fp2::add(g, b, f);
  fp2_half(g, g);
  fp2::add(h, t.y, t.z);
  fp2::sqr(h, h);
  fp2::add(s, b, c);
  fp2::sub(h, h, s);  // h = 2yz
  fp2::sub(i, e, b);
  fp2::sqr(j, t.x);
  fp2::sqr(e2, e);
  fp2::sub(s, b, f);
  fp2::mul(t.x, a, s);
  fp2::sqr(g, g);
  fp2::dbl(s, e2);
  fp2::add(s, s, e2);
  fp2::sub(t.y, g, s);
  fp2::mul(t.z, b, h);
  l.c0 = i;
  fp2::dbl(s, j);
  fp2::add(l.c1, s, j);
  fp2::neg(l.c2, h);
}

inline GS_HD GS_NOINL void g2_add_step(g2_proj& t, const g2_aff& q, line_coeffs& l) {
  fp2 theta, lambda, c, d, e, f, g, h, s, j;
  fp2::mul(s, q.y, t.z);
  fp2::sub(theta, t.y, s);
  fp2::mul(s, q.x, t.z);
  fp2::sub(lambda, t.x, s);
  fp2::sqr(c, theta);
  fp2::sqr(d, lambda);
  fp2::mul(e, lambda, d);
  fp2::mul(f, t.z, c);
  fp2::mul(g, t.x, d);
  fp2::add(h, e, f);
  fp2::sub(h, h, g);
  fp2::sub(h, h, g);
  fp2::mul(t.x, lambda, h);
  fp2::sub(s, g, h);
  fp2::mul(s, theta, s);
  fp2::mul(j, e, t.y);
  fp2::sub(t.y, s, j);
  fp2::mul(t.z, t.z, e);
  fp2::mul(j, theta, q.x);
  fp2::mul(s, lambda, q.y);
  fp2::sub(l.c0, j, s);
  fp2::neg(l.c1, theta);
  l.c2 = lambda;
}

// ------------------------------------------------------------------ affine line walk, E points per thread
// One Miller step (doubling: T <- 2T, or addition: T <- T + Q) for E independent G2 points in AFFINE
// coordinates.  The E slope denominators (2 y_T, or x_T - x_Q) are inverted together: Fp2 norms, Montgomery's
// trick over the E norms, ONE safegcd inversion (modinv.cuh), so a step costs ~13 + 6 Fp products per point
// plus 1/E of an inversion that runs on the integer-ALU pipe.  Output per point: the line through T with
// slope lam on the twist,
//     l(P) * w^3 = yP w^3 - lam xP w^2 + mu,      mu = lam x_T - y_T
// which the caller scales by 1/yP (an Fp factor, killed by the final exponentiation) so that the w^3
// coefficient is ONE:  l' = w^3 + (lam * (-xP/yP)) w^2 + mu * (1/yP).
// Points with act[i] == false are left untouched.  Inputs must be points of the order-r subgroup (as the
// reference's G2Affine values are): then no denominator vanishes; a zero denominator yields lam = 0.
// T / Q are accessor objects (ld(i, x, y), st(i, x, y)): strided shared memory on the GPU, plain arrays in
// tests/hostsim; emit(i, lam, mu) receives each line as soon as it is known (no per-thread line arrays);
// sync() is a block barrier on the GPU (keeps the warps of a block in the same code region, which is what
// lets them share instruction-cache lines) and a no-op in the tests.
// The walk is latency- and I-cache-sensitive (few warps per SM, every warp at its own place in a long loop
// body), so the field products are single out-of-line copies: the loop body stays ~3k instructions.
// (operands and result by value: they stay in registers across the call, see FpOps::mul_v)
inline GS_HD GS_NOINL fp fp_mul_v(fp a, fp b) {
  fp r;
  fp::mul(r, a, b);
  return r;
}
GS_HD GS_INL void fp_mul_n(fp& r, const fp& a, const fp& b) { r = fp_mul_v(a, b); }
inline GS_HD GS_NOINL void fp2_mul_n(fp2& r, const fp2& a, const fp2& b) {
  fp t0, t1, t2, s0, s1;
  fp::add(s0, a.c0, a.c1);
  fp::add(s1, b.c0, b.c1);
  fp_mul_n(t0, a.c0, b.c0);
  fp_mul_n(t1, a.c1, b.c1);
  fp_mul_n(t2, s0, s1);
  fp::sub(r.c0, t0, t1);
  fp::sub(t2, t2, t0);
  fp::sub(r.c1, t2, t1);
}
inline GS_HD GS_NOINL void fp2_sqr_n(fp2& r, const fp2& a) {
  fp s, d, m;
  fp::add(s, a.c0, a.c1);
  fp::sub(d, a.c0, a.c1);
  fp_mul_n(m, a.c0, a.c1);
  fp_mul_n(r.c0, s, d);
  fp::add(r.c1, m, m);
}
inline GS_HD GS_NOINL void fp2_sub_n(fp2& r, const fp2& a, const fp2& b) { fp2::sub(r, a, b); }
inline GS_HD GS_NOINL void fp2_add_n(fp2& r, const fp2& a, const fp2& b) { fp2::add(r, a, b); }

// phase 1 for point i: slope denominator -> its Fp2 norm (1 when the point is inactive or degenerate)
template <class TS, class QS>
GS_HD GS_INL void g2_affine_den(fp2& den, fp2& x, fp2& y, fp2& qx, fp2& qy, TS& T, const QS& Q, int i, bool is_add) {
  T.ld(i, x, y);
  if (is_add) {
    Q.ld(i, qx, qy);
    fp2_sub_n(den, x, qx);
  } else {
    fp2_add_n(den, y, y);
  }
}
template <int E, class TS, class QS, class EM, class SY>
GS_HD GS_INL void g2_affine_step(TS& T, const QS& Q, const bool* act, bool is_add, EM&& emit, SY&& sync) {
  fp nrm[E], pre[E];
#pragma unroll 1
  for (int i = 0; i < E; i++) {
    fp n;
    fp_one(n);
    if (act[i]) {
      fp2 x, y, qx, qy, den;
      g2_affine_den(den, x, y, qx, qy, T, Q, i, is_add);
      fp t;
      fp_mul_n(n, den.c0, den.c0);
      fp_mul_n(t, den.c1, den.c1);
      fp::add(n, n, t);
      if (n.is_zero()) fp_one(n);
    }
    nrm[i] = n;
    if (i == 0)
      pre[0] = n;
    else
      fp_mul_n(pre[i], pre[i - 1], n);
  }
  sync();
  fp inv;
  fp_inv(inv, pre[E - 1]);
  sync();
#pragma unroll 1
  for (int i = E - 1; i >= 0; i--) {
    fp ninv;
    if (i > 0) {
      fp_mul_n(ninv, inv, pre[i - 1]);
      fp_mul_n(inv, inv, nrm[i]);
    } else {
      ninv = inv;
    }
    if (!act[i]) continue;
    fp2 x, y, qx, qy, num, sum, dinv, l, t, x3;
    g2_affine_den(dinv, x, y, qx, qy, T, Q, i, is_add);
    if (is_add) {
      fp2_sub_n(num, y, qy);
      fp2_add_n(sum, x, qx);
    } else {
      fp2_sqr_n(t, x);
      fp2_add_n(num, t, t);
      fp2_add_n(num, num, t);
      fp2_add_n(sum, x, x);
    }
    // 1/den = conj(den) / |den|^2
    fp_mul_n(dinv.c0, dinv.c0, ninv);
    fp_mul_n(dinv.c1, dinv.c1, ninv);
    fp::neg(dinv.c1, dinv.c1);
    fp2_mul_n(l, num, dinv);
    fp2_sqr_n(x3, l);
    fp2_sub_n(x3, x3, sum);
    fp2_mul_n(t, l, x);
    fp2_sub_n(t, t, y);  // mu = lam x_T - y_T
    emit(i, l, t);
    fp2_mul_n(num, l, x3);
    fp2_sub_n(y, t, num);
    T.st(i, x3, y);
  }
}
// array-backed accessor (host tests)
struct g2_pts_arr {
  fp2 *x, *y;
  GS_HD GS_INL void ld(int i, fp2& X, fp2& Y) const {
    X = x[i];
    Y = y[i];
  }
  GS_HD GS_INL void st(int i, const fp2& X, const fp2& Y) {
    x[i] = X;
    y[i] = Y;
  }
};

// Line storage: word-interleaved so that both the writer (one thread per G2 point) and the reader
// (one thread per accumulator) are perfectly coalesced over consecutive problems:
//     word w (0..71) of line `idx` of a point lives at  base[(idx*72 + w) * stride]
constexpr int GS_LINE_WORDS = 72;
GS_HD GS_INL void st_line(uint32_t* base, size_t stride, int idx, const line_coeffs& l) {
  const uint32_t* w = (const uint32_t*)&l;
  uint32_t* o = base + (size_t)idx * GS_LINE_WORDS * stride;
  for (int i = 0; i < GS_LINE_WORDS; i++) o[(size_t)i * stride] = w[i];
}
GS_HD GS_INL void ld_line(line_coeffs& l, const uint32_t* base, size_t stride, int idx) {
  uint32_t* w = (uint32_t*)&l;
  const uint32_t* o = base + (size_t)idx * GS_LINE_WORDS * stride;
  for (int i = 0; i < GS_LINE_WORDS; i++) w[i] = o[(size_t)i * stride];
}

// Writes the 68 line triples of Q (layout above).  Q must not be the identity (callers drop
// identity pairs, as ark-ec does).
GS_HD GS_INL void g2_prepare(uint32_t* out, size_t stride, const g2_aff& q) {
  g2_proj t;
  t.x = q.x;
  t.y = q.y;
  t.z.set_one();
  int idx = 0;
  for (int b = 62; b >= 0; b--) {
    line_coeffs l;
    g2_double_step(t, l);
    st_line(out, stride, idx, l);
    idx++;
    if ((GS_X_ABS >> b) & 1) {
      g2_add_step(t, q, l);
      st_line(out, stride, idx, l);
      idx++;
    }
  }
}

// f *= line evaluated at the affine G1 point (px, py)
GS_HD GS_INL void miller_apply_line(fp12& f, const line_coeffs& l, const fp& px, const fp& py) {
  fp2 c1, c2;
  fp2::mul_fp(c1, l.c1, px);
  fp2::mul_fp(c2, l.c2, py);
  fp12::mul_by_014(f, f, l.c0, c1, c2);
}

// ------------------------------------------------------------------ final exponentiation
// f^|x| for f in the cyclotomic subgroup, then conjugate (x < 0)
inline GS_HD GS_NOINL void cyclotomic_exp_x(fp12& r, const fp12& a) {
  fp12 acc = a;
  for (int b = 62; b >= 0; b--) {
    fp12::cyclotomic_sqr(acc, acc);
    if ((GS_X_ABS >> b) & 1) fp12::mul(acc, acc, a);
  }
  fp12::conj(r, acc);
}

inline GS_HD GS_NOINL void final_exponentiation(fp12& out, const fp12& f) {
  fp12 r, t, a, b, c;
  // easy part: r = f^((p^6-1)(p^2+1))
  fp12::inv(t, f);
  fp12::conj(r, f);
  fp12::mul(r, r, t);
  fp12::frobenius<2>(t, r);
  fp12::mul(r, r, t);
  // hard part: r^((x-1)^2 (x+p)(x^2+p^2-1)) * r^3
  cyclotomic_exp_x(a, r);
  fp12::conj(t, r);
  fp12::mul(a, a, t);  // a = r^(x-1)
  cyclotomic_exp_x(b, a);
  fp12::conj(t, a);
  fp12::mul(b, b, t);  // b = a^(x-1)
  cyclotomic_exp_x(c, b);
  fp12::frobenius<1>(t, b);
  fp12::mul(c, c, t);  // c = b^(x+p)
  cyclotomic_exp_x(a, c);
  cyclotomic_exp_x(a, a);  // c^(x^2)
  fp12::frobenius<2>(t, c);
  fp12::mul(a, a, t);
  fp12::conj(t, c);
  fp12::mul(a, a, t);  // a = c^(x^2+p^2-1)
  fp12::cyclotomic_sqr(t, r);
  fp12::mul(t, t, r);  // r^3
  fp12::mul(out, a, t);
}

}  // namespace gs
