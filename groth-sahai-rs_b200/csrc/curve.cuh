// G1 / G2 group arithmetic for BLS12-381 (y^2 = x^3 + 4 over Fp; M-twist y^2 = x^3 + 4(1+u)
// over Fp2), Jacobian coordinates, templated over the coordinate field.
//
// Replaces: ark-ec `into_group`, `*=`, `+`, `into_affine` at the reference call sites
// src/data_structures.rs:187-188, 336-342, 381-387 and src/generator.rs:57-58, 96-99.
// Outputs are normalised to affine, which is canonical, so the choice of formulas
// cannot influence parity.
//
// Conventions: affine identity = (0, 0) (not on either curve); Jacobian identity = Z == 0.
#pragma once
#include "tower.cuh"

namespace gs {

template <class F>
struct Aff {
  typename F::T x, y;
  GS_HD GS_INL bool is_inf() const { return x.is_zero() && y.is_zero(); }
  GS_HD GS_INL void set_inf() {
    x.set_zero();
    y.set_zero();
  }
};

template <class F>
struct Jac {
  typedef typename F::T T;
  T X, Y, Z;

  GS_HD GS_INL bool is_inf() const { return Z.is_zero(); }
  GS_HD GS_INL void set_inf() {
    F::set_one(X);
    F::set_one(Y);
    Z.set_zero();
  }
  GS_HD GS_INL void from_affine(const Aff<F>& p) {
    if (p.is_inf()) {
      set_inf();
    } else {
      X = p.x;
      Y = p.y;
      F::set_one(Z);
    }
  }

  // dbl-2009-l (a = 0): 2M + 5S
  GS_HD static GS_NOINL void dbl(Jac& r, const Jac& p) {
    T A, B, C, D, E, Fq, t;
    F::sqr(A, p.X);
    F::sqr(B, p.Y);
    F::sqr(C, B);
    F::add(t, p.X, B);
    F::sqr(t, t);
    F::sub(t, t, A);
    F::sub(t, t, C);
    F::dbl(D, t);
    F::dbl(E, A);
    F::add(E, E, A);
    F::sqr(Fq, E);
    F::mul(t, p.Y, p.Z);
    F::dbl(r.Z, t);
    F::sub(t, Fq, D);
    F::sub(r.X, t, D);
    F::sub(t, D, r.X);
    F::mul(t, E, t);
    F::dbl(C, C);
    F::dbl(C, C);
    F::dbl(C, C);
    F::sub(r.Y, t, C);
  }

  // madd-2007-bl: 7M + 4S, all exceptional cases handled explicitly
  GS_HD static GS_NOINL void add_mixed(Jac& r, const Jac& p, const Aff<F>& q) {
    if (q.is_inf()) {
      r = p;
      return;
    }
    if (p.is_inf()) {
      r.X = q.x;
      r.Y = q.y;
      F::set_one(r.Z);
      return;
    }
    T Z1Z1, U2, S2, H, HH, I, J, rr, V, t;
    F::sqr(Z1Z1, p.Z);
    F::mul(U2, q.x, Z1Z1);
    F::mul(S2, q.y, p.Z);
    F::mul(S2, S2, Z1Z1);
    F::sub(H, U2, p.X);
    F::sub(rr, S2, p.Y);
    if (H.is_zero()) {
      if (rr.is_zero()) {
        dbl(r, p);
      } else {
        r.set_inf();
      }
      return;
    }
    F::dbl(rr, rr);
    F::sqr(HH, H);
    F::dbl(I, HH);
    F::dbl(I, I);
    F::mul(J, H, I);
    F::mul(V, p.X, I);
    // Z3 = (Z1 + H)^2 - Z1Z1 - HH
    F::add(t, p.Z, H);
    F::sqr(t, t);
    F::sub(t, t, Z1Z1);
    T Z3;
    F::sub(Z3, t, HH);
    // X3 = r^2 - J - 2V
    F::sqr(t, rr);
    F::sub(t, t, J);
    F::sub(t, t, V);
    T X3;
    F::sub(X3, t, V);
    // Y3 = r (V - X3) - 2 Y1 J
    F::sub(t, V, X3);
    F::mul(t, rr, t);
    F::mul(J, p.Y, J);
    F::dbl(J, J);
    F::sub(r.Y, t, J);
    r.X = X3;
    r.Z = Z3;
  }

  // add-2007-bl: 11M + 5S
  GS_HD static GS_NOINL void add(Jac& r, const Jac& p, const Jac& q) {
    if (q.is_inf()) {
      r = p;
      return;
    }
    if (p.is_inf()) {
      r = q;
      return;
    }
    T Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t;
    F::sqr(Z1Z1, p.Z);
    F::sqr(Z2Z2, q.Z);
    F::mul(U1, p.X, Z2Z2);
    F::mul(U2, q.X, Z1Z1);
    F::mul(S1, p.Y, q.Z);
    F::mul(S1, S1, Z2Z2);
    F::mul(S2, q.Y, p.Z);
    F::mul(S2, S2, Z1Z1);
    F::sub(H, U2, U1);
    F::sub(rr, S2, S1);
    if (H.is_zero()) {
      if (rr.is_zero()) {
        dbl(r, p);
      } else {
        r.set_inf();
      }
      return;
    }
    F::dbl(rr, rr);
    F::dbl(I, H);
    F::sqr(I, I);
    F::mul(J, H, I);
    F::mul(V, U1, I);
    // Z3 = ((Z1+Z2)^2 - Z1Z1 - Z2Z2) H
    F::add(t, p.Z, q.Z);
    F::sqr(t, t);
    F::sub(t, t, Z1Z1);
    F::sub(t, t, Z2Z2);
    T Z3;
    F::mul(Z3, t, H);
    F::sqr(t, rr);
    F::sub(t, t, J);
    F::sub(t, t, V);
    T X3;
    F::sub(X3, t, V);
    F::sub(t, V, X3);
    F::mul(t, rr, t);
    F::mul(J, S1, J);
    F::dbl(J, J);
    F::sub(r.Y, t, J);
    r.X = X3;
    r.Z = Z3;
  }

  GS_HD static GS_INL void neg(Jac& r, const Jac& p) {
    r.X = p.X;
    F::neg(r.Y, p.Y);
    r.Z = p.Z;
  }

  // to affine given zinv = 1/Z (caller batches the inversion); identity -> (0,0)
  GS_HD static GS_NOINL void to_affine_with_zinv(Aff<F>& r, const Jac& p, const T& zinv) {
    if (p.is_inf()) {
      r.set_inf();
      return;
    }
    T z2, z3;
    F::sqr(z2, zinv);
    F::mul(z3, z2, zinv);
    F::mul(r.x, p.X, z2);
    F::mul(r.y, p.Y, z3);
  }
  GS_HD static GS_NOINL void to_affine(Aff<F>& r, const Jac& p) {
    if (p.is_inf()) {
      r.set_inf();
      return;
    }
    T zi;
    F::inv(zi, p.Z);
    to_affine_with_zinv(r, p, zi);
  }
};

typedef Aff<FpOps> g1_aff;
typedef Jac<FpOps> g1_jac;
typedef Aff<Fp2Ops> g2_aff;
typedef Jac<Fp2Ops> g2_jac;

// ------------------------------------------------------------------ scalars
// Montgomery Fr (as stored by arkworks) -> canonical integer limbs
GS_HD GS_INL void fr_from_mont(uint32_t out[8], const fr& a) {
  fr one_raw, t;
  one_raw.set_zero();
  one_raw.l[0] = 1;
  fr::mul(t, a, one_raw);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = t.l[i];
}

// r = k * p, k a canonical 256-bit integer (< r), 4-bit fixed windows, variable-base.
// 252 doublings + <= 64 mixed... (table kept Jacobian: 14 additions to build, no inversion)
template <class F>
GS_HD GS_NOINL void scalar_mul(Jac<F>& r, const Aff<F>& p, const uint32_t k[8]) {
  Jac<F> acc;
  acc.set_inf();
  if (p.is_inf()) {
    r = acc;
    return;
  }
  Jac<F> tab[8];  // 1P .. 8P  (signed digits in [-8, 8))
  tab[0].from_affine(p);
  Jac<F>::dbl(tab[1], tab[0]);
  for (int i = 2; i < 8; i++) Jac<F>::add_mixed(tab[i], tab[i - 1], p);
  // signed 4-bit recoding, MSB first: digits d_i in [-8, 8), carry into the next nibble
  int8_t dig[65];
  int carry = 0;
  for (int i = 0; i < 64; i++) {
    int d = (int)((k[i >> 3] >> ((i & 7) * 4)) & 15u) + carry;
    if (d >= 8) {
      d -= 16;
      carry = 1;
    } else {
      carry = 0;
    }
    dig[i] = (int8_t)d;
  }
  dig[64] = (int8_t)carry;
  for (int i = 64; i >= 0; i--) {
    if (i != 64) {
      Jac<F>::dbl(acc, acc);
      Jac<F>::dbl(acc, acc);
      Jac<F>::dbl(acc, acc);
      Jac<F>::dbl(acc, acc);
    }
    int d = dig[i];
    if (d > 0) {
      Jac<F>::add(acc, acc, tab[d - 1]);
    } else if (d < 0) {
      Jac<F> n;
      Jac<F>::neg(n, tab[-d - 1]);
      Jac<F>::add(acc, acc, n);
    }
  }
  r = acc;
}

}  // namespace gs
