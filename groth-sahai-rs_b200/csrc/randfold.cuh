// Folding primitives of the randomised batch verification (gs_verify_batch_rand, verify.cu): the weights' digit forms and
// the joint double-and-add walks, as host/device functions so that tests/hostsim runs the same code on the CPU.
#pragma once
#include <cstring>

#include "endo.cuh"

namespace gs {

// The G2 weight is beta = b0 + b1 |x| with b0, b1 the two 32-bit halves of the caller's word (2^64 distinct values mod r,
// which is all the soundness argument needs): |x| Y = -psi(Y) is two Fp2 products (endo.cuh), so beta Y.0 is a JOINT
// 33-step double-and-add over (Y.0, -psi(Y.0)) -- half the doublings of a 64-bit scalar -- and in joint sparse form
// (Solinas) only every second step adds.  beta is the same for all threads: the digits are kernel parameters and the
// instruction stream is uniform.
struct jsf33 {
  int8_t u0[34], u1[34];  // digits in {-1, 0, 1}, least significant first
  int len;
};
static inline jsf33 make_jsf(uint32_t a, uint32_t b) {
  jsf33 r;
  memset(&r, 0, sizeof r);
  uint64_t k0 = a, k1 = b;
  int d0 = 0, d1 = 0, n = 0;
  auto digit = [](uint64_t l, uint64_t lo) {
    if ((l & 1) == 0) return 0;
    int u = 2 - (int)(l & 3);
    if (((l & 7) == 3 || (l & 7) == 5) && (lo & 3) == 2) u = -u;
    return u;
  };
  while (k0 + d0 > 0 || k1 + d1 > 0) {
    const uint64_t l0 = k0 + d0, l1 = k1 + d1;
    const int u0 = digit(l0, l1), u1 = digit(l1, l0);
    if (2 * d0 == 1 + u0) d0 = 1 - d0;
    if (2 * d1 == 1 + u1) d1 = 1 - d1;
    k0 >>= 1;
    k1 >>= 1;
    r.u0[n] = (int8_t)u0;
    r.u1[n] = (int8_t)u1;
    n++;
  }
  r.len = n;
  return r;
}
// acc += beta * y0 given p1 = -psi(y0) and the affine sums as = y0 + p1, ad = y0 - p1 (acc must be the identity on entry)
GS_HD GS_INL void rand_fold_g2_walk(g2_jac& acc, const g2_aff& y0, const g2_aff& p1, const g2_aff& as, const g2_aff& ad,
                                         const jsf33& b) {
#pragma unroll 1
  for (int i = b.len - 1; i >= 0; i--) {
    g2_jac::dbl(acc, acc);
    const int u0 = b.u0[i], u1 = b.u1[i];
    if (u0 == 0 && u1 == 0) continue;
    g2_aff t = u1 == 0 ? y0 : (u0 == 0 ? p1 : (u0 == u1 ? as : ad));
    const bool neg = u0 != 0 ? u0 < 0 : u1 < 0;  // the table holds the combinations whose first non-zero digit is +1
    if (neg) fp2::neg(t.y, t.y);
    g2_jac::add_mixed(acc, acc, t);
  }
}
// beta y0 + y1 in one thread (the CRS elements of a call; tests): sums made affine with two field inversions
GS_HD GS_INL void rand_fold_g2_single(g2_aff& out, const g2_aff& y0, const g2_aff& y1, const jsf33& beta) {
  g2_jac acc;
  acc.set_inf();
  if (!y0.is_inf()) {
    g2_aff p1, np1, as, ad;
    endo_psi(p1, y0);
    fp2::neg(p1.y, p1.y);  // -psi(y0) = |x| y0
    np1 = p1;
    fp2::neg(np1.y, p1.y);
    g2_jac js, jd;
    js.from_affine(y0);
    jd = js;
    g2_jac::add_mixed(js, js, p1);
    g2_jac::add_mixed(jd, jd, np1);
    g2_jac::to_affine(as, js);
    g2_jac::to_affine(ad, jd);
    rand_fold_g2_walk(acc, y0, p1, as, ad, beta);
  }
  g2_jac::add_mixed(acc, acc, y1);
  g2_jac::to_affine(out, acc);
}
// acc = sigma tab[1] + tau tab[2] over the table 0, x0, x1, x0 + x1: one table addition per bit, the same instruction
// stream whatever the bits are
GS_HD GS_INL void rand_fold_g1_walk(g1_jac& acc, const g1_aff (&tab)[4], uint64_t sg, uint64_t tu) {
  acc.set_inf();
#pragma unroll 1
  for (int bit = 63; bit >= 0; bit--) {
    g1_jac::dbl(acc, acc);
    const int d = (int)((sg >> bit) & 1) | ((int)((tu >> bit) & 1) << 1);
    g1_jac::add_mixed(acc, acc, tab[d]);
  }
}
GS_HD GS_INL void rand_fold_g1_single(g1_aff& out, const g1_aff& x0, const g1_aff& x1, uint64_t sg, uint64_t tu) {
  g1_aff tab[4];
  tab[0].set_inf();
  tab[1] = x0;
  tab[2] = x1;
  g1_jac j;
  j.from_affine(x0);
  g1_jac::add_mixed(j, j, x1);
  g1_jac::to_affine(tab[3], j);
  g1_jac acc;
  rand_fold_g1_walk(acc, tab, sg, tu);
  g1_jac::to_affine(out, acc);
}

}  // namespace gs
