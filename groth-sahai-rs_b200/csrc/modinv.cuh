// Fp inversion by Bernstein-Yang "safegcd" division steps (constant iteration count, branch free), the
// layout of libsecp256k1's modinv32 carried over to the 381-bit BLS12-381 modulus: 13 signed 30-bit limbs,
// 37 batches of 30 divsteps (1110 >= the proven bound floor((49*381 + 57)/17) = 1101 for delta = 1), each
// batch = 30 steps on the low words (plain 32-bit ALU work) + one 2x2 transition-matrix update of (f, g) and
// (d, e) (26 limbs x 4..6 signed 32x32->64 multiply-adds).
//
// Why: a Fermat inversion a^(p-2) is ~490 Montgomery products = 147k IMAD.WIDE on the pipe that bounds every
// kernel of this library; this is ~5k multiply-adds plus ~35k ALU instructions, which issue on the otherwise
// idle integer-ALU pipe.  It is what makes per-step inversions (affine G2 line walk) affordable.
//
// Replaces: ark-ff `Field::inverse` wherever the reference normalises a point or inverts in the tower
// (src/data_structures.rs:187-188, 336-342 via into_affine; final_exponentiation's easy part).
#pragma once
#include "constants.cuh"

namespace gs {

struct s30 {
  int32_t v[13];
};
constexpr int32_t S30_M = 0x3FFFFFFF;

// canonical 12x32 -> 13x30
GS_HD GS_INL void s30_from_fp(s30& r, const fp& a) {
#pragma unroll
  for (int i = 0; i < 13; i++) {
    const int o = 30 * i, w = o >> 5, s = o & 31;
    uint32_t lo = a.l[w] >> s;
    if (s > 2 && w + 1 < 12) lo |= a.l[w + 1] << (32 - s);
    r.v[i] = (int32_t)(lo & (uint32_t)S30_M);
  }
}
// 13x30 (value in [0, p), limbs in [0, 2^30)) -> 12x32
GS_HD GS_INL void s30_to_fp(fp& r, const s30& a) {
#pragma unroll
  for (int w = 0; w < 12; w++) {
    const int o = 32 * w, i = o / 30, s = o - 30 * i;
    uint32_t x = (uint32_t)a.v[i] >> s;
    x |= (uint32_t)a.v[i + 1] << (30 - s);
    r.l[w] = x;
  }
}

// 30 division steps on the low words; returns the new eta = -delta and the transition matrix t = (u v; q r)
// with  t * (f, g) = 2^30 * (f', g').
GS_HD GS_INL int32_t s30_divsteps(int32_t eta, uint32_t f, uint32_t g, int32_t (&t)[4]) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
#pragma unroll 6
  for (int i = 0; i < 30; i++) {
    uint32_t c1 = (uint32_t)(eta >> 31);  // delta > 0
    uint32_t c2 = 0u - (g & 1u);          // g odd
    uint32_t x = (f ^ c1) - c1, y = (u ^ c1) - c1, z = (v ^ c1) - c1;
    g += x & c2;
    q += y & c2;
    r += z & c2;
    c1 &= c2;  // swap
    eta = (int32_t)(((uint32_t)eta ^ c1) - (c1 + 1u));
    f += g & c1;
    u += q & c1;
    v += r & c1;
    g >>= 1;
    u <<= 1;
    v <<= 1;
  }
  t[0] = (int32_t)u;
  t[1] = (int32_t)v;
  t[2] = (int32_t)q;
  t[3] = (int32_t)r;
  return eta;
}

// (d, e) <- t * (d, e) / 2^30  mod p   (multiples of p added so that the division is exact)
GS_HD GS_INL void s30_update_de(s30& d, s30& e, const int32_t (&t)[4]) {
  const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
  int32_t sd = d.v[12] >> 31, se = e.v[12] >> 31;
  int32_t md = (u & sd) + (v & se);
  int32_t me = (q & sd) + (r & se);
  int32_t di = d.v[0], ei = e.v[0];
  int64_t cd = (int64_t)u * di + (int64_t)v * ei;
  int64_t ce = (int64_t)q * di + (int64_t)r * ei;
  md -= (int32_t)((FP_MODINV30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)S30_M);
  me -= (int32_t)((FP_MODINV30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)S30_M);
  cd += (int64_t)FP_MOD30(0) * md;
  ce += (int64_t)FP_MOD30(0) * me;
  cd >>= 30;
  ce >>= 30;
#pragma unroll
  for (int i = 1; i < 13; i++) {
    di = d.v[i];
    ei = e.v[i];
    cd += (int64_t)u * di + (int64_t)v * ei;
    ce += (int64_t)q * di + (int64_t)r * ei;
    cd += (int64_t)FP_MOD30(i) * md;
    ce += (int64_t)FP_MOD30(i) * me;
    d.v[i - 1] = (int32_t)cd & S30_M;
    cd >>= 30;
    e.v[i - 1] = (int32_t)ce & S30_M;
    ce >>= 30;
  }
  d.v[12] = (int32_t)cd;
  e.v[12] = (int32_t)ce;
}

// (f, g) <- t * (f, g) / 2^30   (exact)
GS_HD GS_INL void s30_update_fg(s30& f, s30& g, const int32_t (&t)[4]) {
  const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
  int32_t fi = f.v[0], gi = g.v[0];
  int64_t cf = (int64_t)u * fi + (int64_t)v * gi;
  int64_t cg = (int64_t)q * fi + (int64_t)r * gi;
  cf >>= 30;
  cg >>= 30;
#pragma unroll
  for (int i = 1; i < 13; i++) {
    fi = f.v[i];
    gi = g.v[i];
    cf += (int64_t)u * fi + (int64_t)v * gi;
    cg += (int64_t)q * fi + (int64_t)r * gi;
    f.v[i - 1] = (int32_t)cf & S30_M;
    cf >>= 30;
    g.v[i - 1] = (int32_t)cg & S30_M;
    cg >>= 30;
  }
  f.v[12] = (int32_t)cf;
  g.v[12] = (int32_t)cg;
}

// r in (-2p, p) -> [0, p), negated first when sign < 0
GS_HD GS_INL void s30_normalize(s30& r, int32_t sign) {
  int32_t cond_add = r.v[12] >> 31;
  int32_t cond_neg = sign >> 31;
#pragma unroll
  for (int i = 0; i < 13; i++) {
    int32_t x = r.v[i] + (FP_MOD30(i) & cond_add);
    r.v[i] = (x ^ cond_neg) - cond_neg;
  }
#pragma unroll
  for (int i = 0; i < 12; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= S30_M;
  }
  cond_add = r.v[12] >> 31;
#pragma unroll
  for (int i = 0; i < 13; i++) r.v[i] += FP_MOD30(i) & cond_add;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    r.v[i + 1] += r.v[i] >> 30;
    r.v[i] &= S30_M;
  }
}

// r = x^-1 mod p for the PLAIN integer x in [0, p) held in the limbs of `a` (0 -> 0)
GS_HD GS_INL void fp_inv_plain(fp& r, const fp& a) {
  s30 d, e, f, g;
#pragma unroll
  for (int i = 0; i < 13; i++) {
    d.v[i] = 0;
    e.v[i] = 0;
    f.v[i] = FP_MOD30(i);
  }
  e.v[0] = 1;
  s30_from_fp(g, a);
  int32_t eta = -1;
#pragma unroll 1
  for (int it = 0; it < 37; it++) {
    int32_t t[4];
    eta = s30_divsteps(eta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
    s30_update_de(d, e, t);
    s30_update_fg(f, g, t);
#if defined(__CUDA_ARCH__)
    // g == 0 in every active lane of the warp: done (further division steps leave d unchanged); typical
    // inputs need ~850 of the 1110 worst-case steps
    int32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 13; i++) nz |= g.v[i];
    if (__all_sync(__activemask(), nz == 0)) break;
#endif
  }
  s30_normalize(d, f.v[12]);
  s30_to_fp(r, d);
}

// Montgomery inverse: a = xR  ->  x^-1 R.   plain_inv(xR) = x^-1 R^-1;  mont_mul(., R^3) = x^-1 R.   0 -> 0.
inline GS_HD GS_NOINL void fp_inv_sg(fp& r, const fp& a) {
  fp t, r3;
  fp_inv_plain(t, a);
#pragma unroll
  for (int i = 0; i < 12; i++) r3.l[i] = FP_R3(i);
  fp::mul(r, t, r3);
}

}  // namespace gs
