// Efficient endomorphisms of BLS12-381 and the scalar splittings built on them.
//   G1:  phi(x, y) = (beta x, y) = -[x^2] P            =>  k P = k1 P + k2 (-phi(P)),  k = k1 + k2 x^2,  k1, k2 < 2^128
//   G2:  psi(x, y) = (conj(x) cx, conj(y) cy) = [x] Q  =>  k Q = sum_j c_j (-1)^j psi^j(Q),  k = sum_j c_j |x|^j,  c_j < 2^64
// (x = -0xd201000000010000 is the curve parameter; both relations hold on the order-r subgroups, which is where every
// G1Affine / G2Affine value of the reference lives: arkworks checks membership when such a value is deserialised, and
// gs_g1_decompress / gs_g2_decompress do the same here.)  The constants were derived from the relations on the
// generators (the serialisation oracle and tests/test_serialize.py).
// Used by: verify.cu (Straus MSM over the GLV halves), prover_impl.cuh (one thread per sub-scalar of an MSM term),
// serial.cu (membership tests).
#pragma once
#include "curve.cuh"

namespace gs {

#define GS_TBL_ENDO_BETA {0x798a64e8u, 0x30f1361bu, 0x7ece5a2au, 0xf3b8ddabu, 0xc61577f7u, 0x16a8ca3au, 0x74fd029bu, 0xc26a2ff8u, 0x60701c6eu, 0x3636b766u, 0x241b6160u, 0x051ba4abu}
#define GS_TBL_PSI_CX_C1 {0x867545c3u, 0x890dc9e4u, 0x3285a5d5u, 0x2af32253u, 0x309b7e2cu, 0x50880866u, 0x7e881024u, 0xa20d1b8cu, 0xe2db9068u, 0x14e4f04fu, 0x1564853au, 0x14e56d3fu}
#define GS_TBL_PSI_CY_C0 {0xa55c9ad1u, 0x3e2f585du, 0x86c18183u, 0x4294213du, 0x8b623732u, 0x382844c8u, 0x19103e18u, 0x92ad2afdu, 0xac7cf0b9u, 0x1d794e4fu, 0x7d825ec8u, 0x0bd592fcu}
#define GS_TBL_PSI_CY_C1 {0x5aa30fdau, 0x7bcfa7a2u, 0x2a927e7cu, 0xdc17dec1u, 0x6b4ebef1u, 0x2f088dd8u, 0xda74d4a7u, 0xd1ca2087u, 0x96cebc1du, 0x2da25966u, 0xbbfd87d2u, 0x0e2b7eedu}
#define GS_TBL_GLV_X2 {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u}               /* x^2 */
#define GS_TBL_GLV_MU {0xf6cfee2eu, 0x63f6e522u, 0xe01faaddu, 0x7c6becf1u, 0x00000001u}  /* 2^256 / x^2 */
// every table exists twice (host array for tests/hostsim, __constant__ array for the kernels), as in constants.cuh
#if defined(__CUDACC__)
#define GS_ENDO_TABLE(name, n)                                        \
  static const uint32_t h_##name[n] = GS_TBL_##name;                  \
  static __device__ __constant__ uint32_t d_##name[n] = GS_TBL_##name;
#else
#define GS_ENDO_TABLE(name, n) static const uint32_t h_##name[n] = GS_TBL_##name;
#endif
#if defined(__CUDA_ARCH__)
#define GS_ENDO_AT(name, i) d_##name[i]
#else
#define GS_ENDO_AT(name, i) h_##name[i]
#endif
GS_ENDO_TABLE(ENDO_BETA, 12)
GS_ENDO_TABLE(PSI_CX_C1, 12)
GS_ENDO_TABLE(PSI_CY_C0, 12)
GS_ENDO_TABLE(PSI_CY_C1, 12)
GS_ENDO_TABLE(GLV_X2, 4)
GS_ENDO_TABLE(GLV_MU, 5)
constexpr uint64_t GS_X_ABS64 = 0xd201000000010000ull;

static GS_HD GS_INL void endo_phi_x(fp& bx, const fp& x) {  // beta * x
  fp beta;
#pragma unroll
  for (int j = 0; j < 12; j++) beta.l[j] = GS_ENDO_AT(ENDO_BETA, j);
  fp::mul(bx, x, beta);
}
// q = psi(p)
static GS_HD GS_NOINL void endo_psi(g2_aff& q, const g2_aff& p) {
  if (p.is_inf()) {
    q = p;
    return;
  }
  fp c1;
  fp2 cy, t;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    c1.l[j] = GS_ENDO_AT(PSI_CX_C1, j);
    cy.c0.l[j] = GS_ENDO_AT(PSI_CY_C0, j);
    cy.c1.l[j] = GS_ENDO_AT(PSI_CY_C1, j);
  }
  // conj(x) * (c1 u) = x.c1 c1 + x.c0 c1 u
  fp a, b;
  fp::mul(a, p.x.c1, c1);
  fp::mul(b, p.x.c0, c1);
  q.x.c0 = a;
  q.x.c1 = b;
  fp2::conj(t, p.y);
  fp2::mul(q.y, t, cy);
}

// k (canonical, < r) -> k1 = k mod x^2, k2 = k div x^2   (Barrett with mu = 2^256 / x^2, at most two corrections)
static GS_HD GS_NOINL void glv_split(uint32_t k1[4], uint32_t k2[4], const uint32_t k[8]) {
  uint32_t t[13];
  for (int i = 0; i < 13; i++) t[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t carry = 0;
    for (int j = 0; j < 5; j++) {
      uint64_t vv = (uint64_t)k[i] * GS_ENDO_AT(GLV_MU, j) + t[i + j] + carry;
      t[i + j] = (uint32_t)vv;
      carry = (uint32_t)(vv >> 32);
    }
    t[i + 5] = carry;
  }
  uint32_t q[4] = {t[8], t[9], t[10], t[11]};  // floor(k mu / 2^256) in {k div x^2 - 1, k div x^2}
  uint32_t pr[8];
  for (int i = 0; i < 8; i++) pr[i] = 0;
  for (int i = 0; i < 4; i++) {
    uint32_t carry = 0;
    for (int j = 0; j < 4; j++) {
      uint64_t vv = (uint64_t)q[i] * GS_ENDO_AT(GLV_X2, j) + pr[i + j] + carry;
      pr[i + j] = (uint32_t)vv;
      carry = (uint32_t)(vv >> 32);
    }
    pr[i + 4] = carry;
  }
  uint32_t rem[8];
  uint32_t borrow = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t d = (uint64_t)k[i] - pr[i] - borrow;
    rem[i] = (uint32_t)d;
    borrow = (uint32_t)(d >> 63);
  }
  for (int it = 0; it < 2; it++) {
    bool ge = (rem[4] | rem[5] | rem[6] | rem[7]) != 0;
    if (!ge) {
      ge = true;
      for (int i = 3; i >= 0; i--)
        if (rem[i] != GS_ENDO_AT(GLV_X2, i)) {
          ge = rem[i] > GS_ENDO_AT(GLV_X2, i);
          break;
        }
    }
    if (!ge) break;
    borrow = 0;
    for (int i = 0; i < 8; i++) {
      uint64_t d = (uint64_t)rem[i] - (i < 4 ? GS_ENDO_AT(GLV_X2, i) : 0u) - borrow;
      rem[i] = (uint32_t)d;
      borrow = (uint32_t)(d >> 63);
    }
    uint32_t c = 1;
    for (int i = 0; i < 4; i++) {
      uint64_t a = (uint64_t)q[i] + c;
      q[i] = (uint32_t)a;
      c = (uint32_t)(a >> 32);
    }
  }
  for (int i = 0; i < 4; i++) {
    k1[i] = rem[i];
    k2[i] = q[i];
  }
}

// k (canonical, < r < |x|^4) -> digits c_0..c_3 in base |x|
static GS_HD GS_NOINL void gls_split(uint64_t c[4], const uint32_t k[8]) {
  uint32_t t[8];
  for (int i = 0; i < 8; i++) t[i] = k[i];
  for (int j = 0; j < 3; j++) {
    unsigned __int128 rem = 0;
    for (int i = 7; i >= 0; i--) {
      unsigned __int128 cur = (rem << 32) | t[i];
      uint64_t q = (uint64_t)(cur / GS_X_ABS64);  // < 2^32 because rem < |x|
      rem = cur - (unsigned __int128)q * GS_X_ABS64;
      t[i] = (uint32_t)q;
    }
    c[j] = (uint64_t)rem;
  }
  c[3] = (uint64_t)t[0] | ((uint64_t)t[1] << 32);
}

// r = k * p for a short scalar: `nwin` signed 4-bit windows over the low 4*nwin bits of k (limbs beyond are ignored)
template <class F>
GS_HD GS_NOINL void scalar_mul_win(Jac<F>& r, const Aff<F>& p, const uint32_t* k, int nwin) {
  Jac<F> acc;
  acc.set_inf();
  if (p.is_inf()) {
    r = acc;
    return;
  }
  Jac<F> tab[8];  // 1P .. 8P
  tab[0].from_affine(p);
  Jac<F>::dbl(tab[1], tab[0]);
  for (int i = 2; i < 8; i++) Jac<F>::add_mixed(tab[i], tab[i - 1], p);
  int8_t dig[65];
  int carry = 0;
  for (int i = 0; i < nwin; i++) {
    int d = (int)((k[i >> 3] >> ((i & 7) * 4)) & 15u) + carry;
    if (d >= 8) {
      d -= 16;
      carry = 1;
    } else {
      carry = 0;
    }
    dig[i] = (int8_t)d;
  }
  dig[nwin] = (int8_t)carry;
  for (int i = nwin; i >= 0; i--) {
    if (i != nwin) {
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

// Sub-scalar j of the endomorphism splitting of (k, base): r = part_j with  k * base = sum_j part_j
template <class F>
struct EndoSplit;
template <>
struct EndoSplit<FpOps> {
  static constexpr int PARTS = 2;
  GS_HD static GS_INL void part(g1_jac& r, const g1_aff& base, const uint32_t k[8], int j) {
    uint32_t k1[4], k2[4];
    glv_split(k1, k2, k);
    g1_aff b = base;
    if (j == 1 && !b.is_inf()) {  // -phi(P) = (beta x, -y)
      endo_phi_x(b.x, base.x);
      fp::neg(b.y, base.y);
    }
    scalar_mul_win<FpOps>(r, b, j == 0 ? k1 : k2, 32);
  }
};
template <>
struct EndoSplit<Fp2Ops> {
  static constexpr int PARTS = 4;
  GS_HD static GS_INL void part(g2_jac& r, const g2_aff& base, const uint32_t k[8], int j) {
    uint64_t c[4];
    gls_split(c, k);
    g2_aff b = base;
    for (int i = 0; i < j; i++) {
      g2_aff t;
      endo_psi(t, b);
      b = t;
    }
    if ((j & 1) && !b.is_inf()) fp2::neg(b.y, b.y);  // |x|^j = (-1)^j x^j
    uint32_t kk[2] = {(uint32_t)c[j], (uint32_t)(c[j] >> 32)};
    scalar_mul_win<Fp2Ops>(r, b, kk, 16);
  }
};
}  // namespace gs
