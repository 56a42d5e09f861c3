/* gs_oracle.c -- CPU restatement (plain C, 6 x 64-bit limbs) of the reference's hot path.
 * TEST INFRASTRUCTURE ONLY: the checker and the CPU baseline, never the product.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * What it restates (paths relative to /root/reference; arithmetic = arkworks 0.5's published
 * algorithms, which are not vendored there -- see oracle/bls12_381.py for the conventions):
 *   src/data_structures.rs:336-342   Com::scalar_mul  = 2 variable-base scalar muls + 2 into_affine
 *   src/data_structures.rs:185-190   Com + Com        = affine add + normalisation per coordinate
 *   src/data_structures.rs:494-502   ComT::pairing_sum = 4 x multi_pairing (each with its own final exp)
 *   src/data_structures.rs:696-742   Mat::left_mul    = term-by-term scalar_mul then Sum
 *   src/verifier.rs:23-55            PPE::verify in the reference's order: 5 pairing_sums (20 final
 *                                    exponentiations), Gamma*d as m*n Com2 scalar muls
 *   src/prover/commit.rs:78-100      batch_commit_G1 (single-threaded, as in the reference); :178-200 G2 mirror
 *   src/prover/commit.rs:125-156     batch_commit_scalar_to_B1 (x W1 + r u1, W1 recomputed per element as
 *                                    data_structures.rs:323-326 does); :225-256 B2 mirror
 *   src/prover/prove.rs:92-171, 195-274, 298-379, 409-488   Provable::prove for PPE / MSMEG1 / MSMEG2 / Quad,
 *                                    every left_mul term by term (one Com::scalar_mul + one affine Com add per
 *                                    term), the Fr products R^T Gamma etc. as Matrix<Fr>::right_mul
 *   src/verifier.rs:23-157           Verifiable::verify for the four types (gsref_verify), same order
 *   src/data_structures.rs:509-540   the four iota_T maps
 * PARITY STATUS: parity unpinned against arkworks bits (no golden vectors exist offline); this file
 * is checked against the independent big-int oracle (tests/test_c_oracle.py).
 *
 * Threading: the reference parallelises only inside left_mul (Rayon over output rows).  For batch
 * workloads the baseline additionally spreads independent proofs over `nthreads` host threads
 * (pthread), which is MORE parallelism than the reference has -- stated wherever it is reported.
 * gsref_prove / gsref_verify spread the scalar multiplications of a left_mul over `nthreads` (term-level,
 * again more than Rayon's <= 2 tasks in prove); results do not depend on it (outputs are canonical affine points).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fp;
typedef struct { fp c0, c1; } fp2;
typedef struct { fp2 c0, c1, c2; } fp6;
typedef struct { fp6 c0, c1; } fp12;
typedef struct { uint64_t l[4]; } fr;

static const uint64_t P[6] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                              0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const uint64_t PINV = 0x89f3fffcfffcfffdull; /* -p^-1 mod 2^64 */
static const fp FP_ONE = {{0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull,
                           0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull}};
static const uint64_t RMOD[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const uint64_t RINV = 0xfffffffeffffffffull; /* -r^-1 mod 2^64 */
#define X_ABS 0xd201000000010000ull

/* ------------------------------------------------------------------ Fp */
static int fp_is_zero(const fp* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3] | a->l[4] | a->l[5]) == 0; }
static int fp_eq(const fp* a, const fp* b) { return memcmp(a, b, sizeof(fp)) == 0; }
static int geq_p(const uint64_t* t) {
  for (int i = 5; i >= 0; i--) {
    if (t[i] > P[i]) return 1;
    if (t[i] < P[i]) return 0;
  }
  return 1;
}
static void sub_p(uint64_t* t) {
  u128 b = 0;
  for (int i = 0; i < 6; i++) {
    u128 d = (u128)t[i] - P[i] - b;
    t[i] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
}
static void fp_add(fp* r, const fp* a, const fp* b) {
  u128 c = 0;
  uint64_t t[6];
  for (int i = 0; i < 6; i++) {
    c += (u128)a->l[i] + b->l[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  if (geq_p(t)) sub_p(t);
  memcpy(r->l, t, 48);
}
static void fp_sub(fp* r, const fp* a, const fp* b) {
  u128 br = 0;
  uint64_t t[6];
  for (int i = 0; i < 6; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - br;
    t[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
  if (br) {
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
      c += (u128)t[i] + P[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  memcpy(r->l, t, 48);
}
static void fp_neg(fp* r, const fp* a) {
  fp z;
  memset(&z, 0, sizeof z);
  fp_sub(r, &z, a);
}
static void fp_mul(fp* r, const fp* a, const fp* b) { /* CIOS Montgomery */
  uint64_t t[8] = {0};
  for (int i = 0; i < 6; i++) {
    u128 c = 0;
    for (int j = 0; j < 6; j++) {
      c += (u128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[6];
    t[6] = (uint64_t)c;
    t[7] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * PINV;
    c = (u128)m * P[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 6; j++) {
      c += (u128)m * P[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[6];
    t[5] = (uint64_t)c;
    t[6] = t[7] + (uint64_t)(c >> 64);
  }
  if (t[6] || geq_p(t)) sub_p(t);
  memcpy(r->l, t, 48);
}
/* dedicated squaring: the 15 cross products once and doubled, then 6 Montgomery reduction rows (what a tuned CPU library
 * does; keeps the CPU baseline honest) */
static void fp_sqr(fp* r, const fp* a) {
  uint64_t t[13] = {0};
  for (int i = 0; i < 5; i++) { /* cross terms a_i a_j, i < j */
    u128 c = 0;
    for (int j = i + 1; j < 6; j++) {
      c += (u128)a->l[i] * a->l[j] + t[i + j];
      t[i + j] = (uint64_t)c;
      c >>= 64;
    }
    t[i + 6] = (uint64_t)c;
  }
  uint64_t top = 0; /* double */
  for (int i = 1; i < 12; i++) {
    uint64_t n = t[i] >> 63;
    t[i] = (t[i] << 1) | top;
    top = n;
  }
  u128 c = 0; /* + squares on the diagonal */
  for (int i = 0; i < 6; i++) {
    u128 sq = (u128)a->l[i] * a->l[i];
    c += (u128)t[2 * i] + (uint64_t)sq;
    t[2 * i] = (uint64_t)c;
    c >>= 64;
    c += (u128)t[2 * i + 1] + (uint64_t)(sq >> 64);
    t[2 * i + 1] = (uint64_t)c;
    c >>= 64;
  }
  for (int i = 0; i < 6; i++) { /* Montgomery reduction */
    uint64_t m = t[i] * PINV;
    u128 d = (u128)m * P[0] + t[i];
    d >>= 64;
    for (int j = 1; j < 6; j++) {
      d += (u128)m * P[j] + t[i + j];
      t[i + j] = (uint64_t)d;
      d >>= 64;
    }
    for (int j = i + 6; d && j < 13; j++) {
      d += t[j];
      t[j] = (uint64_t)d;
      d >>= 64;
    }
  }
  uint64_t o[7];
  memcpy(o, t + 6, 56);
  if (o[6] || geq_p(o)) sub_p(o);
  memcpy(r->l, o, 48);
}
static void fp_inv_fermat(fp* r, const fp* a) { /* a^(p-2), square-and-multiply (kept as a cross-check) */
  uint64_t e[6];
  memcpy(e, P, 48);
  e[0] -= 2;
  fp acc = FP_ONE, base = *a;
  for (int i = 0; i < 384; i++) {
    if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(&acc, &acc, &base);
    fp_sqr(&base, &base);
  }
  *r = acc;
}
/* Binary extended Euclid on the Montgomery residue (the algorithm ark-ff's Fp::inverse uses: Guajardo et al.,
 * alg. 16, started from b = R^2 so that the result is the Montgomery form of the inverse). */
static fp FP_R2;
static pthread_once_t r2_once = PTHREAD_ONCE_INIT;
static void r2_init(void) {
  fp t = FP_ONE; /* R mod p */
  for (int i = 0; i < 384; i++) fp_add(&t, &t, &t);
  FP_R2 = t; /* R * 2^384 = R^2 mod p */
}
static int big6_is_one(const uint64_t* t) { return t[0] == 1 && (t[1] | t[2] | t[3] | t[4] | t[5]) == 0; }
static int big6_geq(const uint64_t* a, const uint64_t* b) {
  for (int i = 5; i >= 0; i--) {
    if (a[i] > b[i]) return 1;
    if (a[i] < b[i]) return 0;
  }
  return 1;
}
static void big6_sub(uint64_t* a, const uint64_t* b) {
  u128 br = 0;
  for (int i = 0; i < 6; i++) {
    u128 d = (u128)a[i] - b[i] - br;
    a[i] = (uint64_t)d;
    br = (d >> 64) & 1;
  }
}
static void big6_shr1(uint64_t* a, uint64_t top) {
  for (int i = 0; i < 5; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 63);
  a[5] = (a[5] >> 1) | (top << 63);
}
static void half_mod_p(uint64_t* b) { /* b/2 mod p */
  uint64_t carry = 0;
  if (b[0] & 1) {
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
      c += (u128)b[i] + P[i];
      b[i] = (uint64_t)c;
      c >>= 64;
    }
    carry = (uint64_t)c;
  }
  big6_shr1(b, carry);
}
static void fp_inv(fp* r, const fp* a) {
  if (fp_is_zero(a)) { memset(r, 0, sizeof *r); return; }
  pthread_once(&r2_once, r2_init);
  uint64_t u[6], v[6];
  fp b = FP_R2, c;
  memset(&c, 0, sizeof c);
  memcpy(u, a->l, 48);
  memcpy(v, P, 48);
  while (!big6_is_one(u) && !big6_is_one(v)) {
    while (!(u[0] & 1)) { big6_shr1(u, 0); half_mod_p(b.l); }
    while (!(v[0] & 1)) { big6_shr1(v, 0); half_mod_p(c.l); }
    if (big6_geq(u, v)) { big6_sub(u, v); fp_sub(&b, &b, &c); }
    else { big6_sub(v, u); fp_sub(&c, &c, &b); }
  }
  *r = big6_is_one(u) ? b : c;
}

/* ------------------------------------------------------------------ Fp2 */
static void fp2_add(fp2* r, const fp2* a, const fp2* b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static void fp2_sub(fp2* r, const fp2* a, const fp2* b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static void fp2_neg(fp2* r, const fp2* a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static void fp2_dbl(fp2* r, const fp2* a) { fp2_add(r, a, a); }
static int fp2_is_zero(const fp2* a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static int fp2_eq(const fp2* a, const fp2* b) { return memcmp(a, b, sizeof(fp2)) == 0; }
static void fp2_mul(fp2* r, const fp2* a, const fp2* b) {
  fp t0, t1, s0, s1, t2;
  fp_mul(&t0, &a->c0, &b->c0);
  fp_mul(&t1, &a->c1, &b->c1);
  fp_add(&s0, &a->c0, &a->c1);
  fp_add(&s1, &b->c0, &b->c1);
  fp_mul(&t2, &s0, &s1);
  fp_sub(&r->c0, &t0, &t1);
  fp_sub(&t2, &t2, &t0);
  fp_sub(&r->c1, &t2, &t1);
}
static void fp2_sqr(fp2* r, const fp2* a) {
  fp s, d, m;
  fp_add(&s, &a->c0, &a->c1);
  fp_sub(&d, &a->c0, &a->c1);
  fp_mul(&m, &a->c0, &a->c1);
  fp_mul(&r->c0, &s, &d);
  fp_add(&r->c1, &m, &m);
}
static void fp2_mul_fp(fp2* r, const fp2* a, const fp* b) { fp_mul(&r->c0, &a->c0, b); fp_mul(&r->c1, &a->c1, b); }
static void fp2_mul_xi(fp2* r, const fp2* a) { /* (1+u) */
  fp t0, t1;
  fp_sub(&t0, &a->c0, &a->c1);
  fp_add(&t1, &a->c0, &a->c1);
  r->c0 = t0;
  r->c1 = t1;
}
static void fp2_inv(fp2* r, const fp2* a) {
  fp n, t;
  fp_sqr(&n, &a->c0);
  fp_sqr(&t, &a->c1);
  fp_add(&n, &n, &t);
  fp_inv(&n, &n);
  fp_mul(&r->c0, &a->c0, &n);
  fp_mul(&t, &a->c1, &n);
  fp_neg(&r->c1, &t);
}
static void fp2_conj(fp2* r, const fp2* a) { r->c0 = a->c0; fp_neg(&r->c1, &a->c1); }

/* ------------------------------------------------------------------ Fp6 / Fp12 */
static void fp6_add(fp6* r, const fp6* a, const fp6* b) { fp2_add(&r->c0, &a->c0, &b->c0); fp2_add(&r->c1, &a->c1, &b->c1); fp2_add(&r->c2, &a->c2, &b->c2); }
static void fp6_sub(fp6* r, const fp6* a, const fp6* b) { fp2_sub(&r->c0, &a->c0, &b->c0); fp2_sub(&r->c1, &a->c1, &b->c1); fp2_sub(&r->c2, &a->c2, &b->c2); }
static void fp6_neg(fp6* r, const fp6* a) { fp2_neg(&r->c0, &a->c0); fp2_neg(&r->c1, &a->c1); fp2_neg(&r->c2, &a->c2); }
static void fp6_mul_v(fp6* r, const fp6* a) {
  fp2 t;
  fp2_mul_xi(&t, &a->c2);
  r->c2 = a->c1;
  r->c1 = a->c0;
  r->c0 = t;
}
static void fp6_mul(fp6* r, const fp6* a, const fp6* b) { /* schoolbook-with-Karatsuba, 6 Fp2 products */
  fp2 v0, v1, v2, t, s, u, c0, c1, c2;
  fp2_mul(&v0, &a->c0, &b->c0);
  fp2_mul(&v1, &a->c1, &b->c1);
  fp2_mul(&v2, &a->c2, &b->c2);
  fp2_add(&t, &a->c1, &a->c2); fp2_add(&s, &b->c1, &b->c2); fp2_mul(&u, &t, &s);
  fp2_sub(&u, &u, &v1); fp2_sub(&u, &u, &v2); fp2_mul_xi(&u, &u); fp2_add(&c0, &u, &v0);
  fp2_add(&t, &a->c0, &a->c1); fp2_add(&s, &b->c0, &b->c1); fp2_mul(&u, &t, &s);
  fp2_sub(&u, &u, &v0); fp2_sub(&u, &u, &v1); fp2_mul_xi(&t, &v2); fp2_add(&c1, &u, &t);
  fp2_add(&t, &a->c0, &a->c2); fp2_add(&s, &b->c0, &b->c2); fp2_mul(&u, &t, &s);
  fp2_sub(&u, &u, &v0); fp2_sub(&u, &u, &v2); fp2_add(&c2, &u, &v1);
  r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static void fp6_inv(fp6* r, const fp6* a) {
  fp2 t0, t1, t2, s, d;
  fp2_sqr(&t0, &a->c0); fp2_mul(&s, &a->c1, &a->c2); fp2_mul_xi(&s, &s); fp2_sub(&t0, &t0, &s);
  fp2_sqr(&t1, &a->c2); fp2_mul_xi(&t1, &t1); fp2_mul(&s, &a->c0, &a->c1); fp2_sub(&t1, &t1, &s);
  fp2_sqr(&t2, &a->c1); fp2_mul(&s, &a->c0, &a->c2); fp2_sub(&t2, &t2, &s);
  fp2_mul(&d, &a->c2, &t1); fp2_mul(&s, &a->c1, &t2); fp2_add(&d, &d, &s); fp2_mul_xi(&d, &d);
  fp2_mul(&s, &a->c0, &t0); fp2_add(&d, &d, &s); fp2_inv(&d, &d);
  fp2_mul(&r->c0, &t0, &d); fp2_mul(&r->c1, &t1, &d); fp2_mul(&r->c2, &t2, &d);
}
static void fp12_one(fp12* r) { memset(r, 0, sizeof *r); r->c0.c0.c0 = FP_ONE; }
static int fp12_eq(const fp12* a, const fp12* b) { return memcmp(a, b, sizeof(fp12)) == 0; }
static void fp12_mul(fp12* r, const fp12* a, const fp12* b) {
  fp6 aa, bb, s, t;
  fp6_mul(&aa, &a->c0, &b->c0);
  fp6_mul(&bb, &a->c1, &b->c1);
  fp6_add(&s, &a->c0, &a->c1);
  fp6_add(&t, &b->c0, &b->c1);
  fp6_mul(&s, &s, &t);
  fp6_sub(&s, &s, &aa);
  fp6_sub(&r->c1, &s, &bb);
  fp6_mul_v(&bb, &bb);
  fp6_add(&r->c0, &aa, &bb);
}
static void fp12_sqr(fp12* r, const fp12* a) { /* complex squaring, 2 Fp6 products */
  fp6 ab, s, t;
  fp6_mul(&ab, &a->c0, &a->c1);
  fp6_add(&s, &a->c0, &a->c1);
  fp6_mul_v(&t, &a->c1);
  fp6_add(&t, &t, &a->c0);
  fp6_mul(&s, &s, &t);
  fp6_sub(&s, &s, &ab);
  fp6_mul_v(&t, &ab);
  fp6_sub(&r->c0, &s, &t);
  fp6_add(&r->c1, &ab, &ab);
}
static void fp4_sq(fp2* lo, fp2* hi, const fp2* a, const fp2* b) { /* (a + b y)^2, y^2 = xi */
  fp2 ab, s, t;
  fp2_mul(&ab, a, b); fp2_add(&s, a, b); fp2_mul_xi(&t, b); fp2_add(&t, &t, a); fp2_mul(&s, &s, &t);
  fp2_sub(&s, &s, &ab); fp2_mul_xi(&t, &ab); fp2_sub(lo, &s, &t); fp2_dbl(hi, &ab);
}
static void fp12_cyclo_sqr(fp12* r, const fp12* a) { /* Granger-Scott */
  fp2 t0, t1, t2, t3, t4, t5, x, z0, z1, z2, z3, z4, z5, tmp;
  fp4_sq(&t0, &t1, &a->c0.c0, &a->c1.c1);
  fp4_sq(&t2, &t3, &a->c1.c0, &a->c0.c2);
  fp4_sq(&t4, &t5, &a->c0.c1, &a->c1.c2);
  fp2_sub(&x, &t0, &a->c0.c0); fp2_dbl(&x, &x); fp2_add(&z0, &x, &t0);
  fp2_add(&x, &t1, &a->c1.c1); fp2_dbl(&x, &x); fp2_add(&z1, &x, &t1);
  fp2_mul_xi(&tmp, &t5); fp2_add(&x, &tmp, &a->c1.c0); fp2_dbl(&x, &x); fp2_add(&z2, &x, &tmp);
  fp2_sub(&x, &t4, &a->c0.c2); fp2_dbl(&x, &x); fp2_add(&z3, &x, &t4);
  fp2_sub(&x, &t2, &a->c0.c1); fp2_dbl(&x, &x); fp2_add(&z4, &x, &t2);
  fp2_add(&x, &t3, &a->c1.c2); fp2_dbl(&x, &x); fp2_add(&z5, &x, &t3);
  r->c0.c0 = z0; r->c1.c1 = z1; r->c1.c0 = z2; r->c0.c2 = z3; r->c0.c1 = z4; r->c1.c2 = z5;
}
static void fp12_conj(fp12* r, const fp12* a) { r->c0 = a->c0; fp6_neg(&r->c1, &a->c1); }
static void fp12_inv(fp12* r, const fp12* a) {
  fp6 t0, t1;
  fp6_mul(&t0, &a->c0, &a->c0);
  fp6_mul(&t1, &a->c1, &a->c1);
  fp6_mul_v(&t1, &t1);
  fp6_sub(&t0, &t0, &t1);
  fp6_inv(&t0, &t0);
  fp6_mul(&r->c0, &a->c0, &t0);
  fp6_mul(&t1, &a->c1, &t0);
  fp6_neg(&r->c1, &t1);
}
/* Frobenius coefficients xi^(i (p^K-1)/6), computed at start-up by exponentiation */
static fp2 FROB[2][6];
static pthread_once_t frob_once = PTHREAD_ONCE_INIT;
static void fp2_pow_big(fp2* r, const fp2* a, const uint64_t* e, int nlimbs) {
  fp2 acc, base = *a;
  memset(&acc, 0, sizeof acc);
  acc.c0 = FP_ONE;
  for (int i = 0; i < nlimbs * 64; i++) {
    if ((e[i >> 6] >> (i & 63)) & 1) fp2_mul(&acc, &acc, &base);
    fp2_sqr(&base, &base);
  }
  *r = acc;
}
static void big_mul(uint64_t* r, const uint64_t* a, int na, const uint64_t* b, int nb) {
  memset(r, 0, (na + nb) * 8);
  for (int i = 0; i < na; i++) {
    u128 c = 0;
    for (int j = 0; j < nb; j++) {
      c += (u128)a[i] * b[j] + r[i + j];
      r[i + j] = (uint64_t)c;
      c >>= 64;
    }
    r[i + nb] = (uint64_t)c;
  }
}
static void big_div_small(uint64_t* a, int n, uint64_t d) {
  u128 rem = 0;
  for (int i = n - 1; i >= 0; i--) {
    u128 cur = (rem << 64) | a[i];
    a[i] = (uint64_t)(cur / d);
    rem = cur % d;
  }
}
static void frob_init(void) {
  fp2 xi;
  xi.c0 = FP_ONE;
  xi.c1 = FP_ONE;
  uint64_t e1[6], e2[12];
  memcpy(e1, P, 48);
  e1[0] -= 1;
  big_div_small(e1, 6, 6); /* (p-1)/6 */
  big_mul(e2, P, 6, P, 6);
  e2[0] -= 1;
  big_div_small(e2, 12, 6); /* (p^2-1)/6 */
  fp2 g1, g2;
  fp2_pow_big(&g1, &xi, e1, 6);
  fp2_pow_big(&g2, &xi, e2, 12);
  for (int K = 0; K < 2; K++) {
    fp2 g = K ? g2 : g1, acc;
    memset(&acc, 0, sizeof acc);
    acc.c0 = FP_ONE;
    for (int i = 0; i < 6; i++) {
      FROB[K][i] = acc;
      fp2_mul(&acc, &acc, &g);
    }
  }
}
static void fp12_frob(fp12* r, const fp12* a, int K) { /* K = 1 or 2 */
  pthread_once(&frob_once, frob_init);
  const fp2* src[6] = {&a->c0.c0, &a->c1.c0, &a->c0.c1, &a->c1.c1, &a->c0.c2, &a->c1.c2};
  fp2 out[6];
  for (int i = 0; i < 6; i++) {
    fp2 t = *src[i];
    if (K & 1) fp2_conj(&t, &t);
    fp2_mul(&out[i], &t, &FROB[K - 1][i]);
  }
  r->c0.c0 = out[0]; r->c1.c0 = out[1]; r->c0.c1 = out[2]; r->c1.c1 = out[3]; r->c0.c2 = out[4]; r->c1.c2 = out[5];
}

/* ------------------------------------------------------------------ curves: affine (0,0) = identity */
typedef struct { fp x, y; } g1a;
typedef struct { fp2 x, y; } g2a;
typedef struct { fp x, y, z; } g1j;
typedef struct { fp2 x, y, z; } g2j;
static int g1a_inf(const g1a* p) { return fp_is_zero(&p->x) && fp_is_zero(&p->y); }
static int g2a_inf(const g2a* p) { return fp2_is_zero(&p->x) && fp2_is_zero(&p->y); }

#define CURVE_IMPL(G, F, FT, ONE_INIT)                                                                          \
  static void G##j_set_inf(G##j* r) { memset(r, 0, sizeof *r); }                                                \
  static int G##j_inf(const G##j* p) { return F##_is_zero(&p->z); }                                             \
  static void G##j_from_affine(G##j* r, const G##a* p) {                                                        \
    if (G##a_inf(p)) { G##j_set_inf(r); return; }                                                               \
    r->x = p->x; r->y = p->y; memset(&r->z, 0, sizeof r->z); ONE_INIT;                                          \
  }                                                                                                             \
  static void G##j_dbl(G##j* r, const G##j* p) {                                                                \
    if (G##j_inf(p)) { *r = *p; return; }                                                                       \
    FT a, b, c, d, e, f, t;                                                                                     \
    F##_sqr(&a, &p->x); F##_sqr(&b, &p->y); F##_sqr(&c, &b);                                                    \
    F##_add(&t, &p->x, &b); F##_sqr(&t, &t); F##_sub(&t, &t, &a); F##_sub(&t, &t, &c); F##_add(&d, &t, &t);      \
    F##_add(&e, &a, &a); F##_add(&e, &e, &a); F##_sqr(&f, &e);                                                  \
    F##_mul(&t, &p->y, &p->z); F##_add(&r->z, &t, &t);                                                          \
    F##_sub(&t, &f, &d); F##_sub(&r->x, &t, &d);                                                                \
    F##_sub(&t, &d, &r->x); F##_mul(&t, &e, &t);                                                                \
    F##_add(&c, &c, &c); F##_add(&c, &c, &c); F##_add(&c, &c, &c); F##_sub(&r->y, &t, &c);                      \
  }                                                                                                             \
  static void G##j_add(G##j* r, const G##j* p, const G##j* q) {                                                 \
    if (G##j_inf(q)) { *r = *p; return; }                                                                       \
    if (G##j_inf(p)) { *r = *q; return; }                                                                       \
    FT z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t, x3, z3;                                                   \
    F##_sqr(&z1z1, &p->z); F##_sqr(&z2z2, &q->z);                                                               \
    F##_mul(&u1, &p->x, &z2z2); F##_mul(&u2, &q->x, &z1z1);                                                     \
    F##_mul(&s1, &p->y, &q->z); F##_mul(&s1, &s1, &z2z2);                                                       \
    F##_mul(&s2, &q->y, &p->z); F##_mul(&s2, &s2, &z1z1);                                                       \
    F##_sub(&h, &u2, &u1); F##_sub(&rr, &s2, &s1);                                                              \
    if (F##_is_zero(&h)) { if (F##_is_zero(&rr)) G##j_dbl(r, p); else G##j_set_inf(r); return; }                \
    F##_add(&rr, &rr, &rr); F##_add(&i, &h, &h); F##_sqr(&i, &i); F##_mul(&j, &h, &i); F##_mul(&v, &u1, &i);    \
    F##_add(&t, &p->z, &q->z); F##_sqr(&t, &t); F##_sub(&t, &t, &z1z1); F##_sub(&t, &t, &z2z2);                 \
    F##_mul(&z3, &t, &h);                                                                                       \
    F##_sqr(&t, &rr); F##_sub(&t, &t, &j); F##_sub(&t, &t, &v); F##_sub(&x3, &t, &v);                           \
    F##_sub(&t, &v, &x3); F##_mul(&t, &rr, &t); F##_mul(&j, &s1, &j); F##_add(&j, &j, &j);                      \
    F##_sub(&r->y, &t, &j); r->x = x3; r->z = z3;                                                               \
  }                                                                                                             \
  static void G##j_to_affine(G##a* r, const G##j* p) { /* into_affine: one inversion */                         \
    if (G##j_inf(p)) { memset(r, 0, sizeof *r); return; }                                                       \
    FT zi, z2, z3;                                                                                              \
    F##_inv(&zi, &p->z); F##_sqr(&z2, &zi); F##_mul(&z3, &z2, &zi);                                             \
    F##_mul(&r->x, &p->x, &z2); F##_mul(&r->y, &p->y, &z3);                                                     \
  }                                                                                                             \
  /* variable-base scalar multiplication (4-bit fixed windows: about the cost of ark-bls12-381's GLV path),  */  \
  /* then into_affine                                                                                         */  \
  static void G##a_mul(G##a* r, const G##a* p, const uint64_t k[4]) {                                           \
    G##j tab[16], acc; G##j_set_inf(&tab[0]); G##j_from_affine(&tab[1], p);                                     \
    for (int i = 2; i < 16; i++) { if (i & 1) G##j_add(&tab[i], &tab[i - 1], &tab[1]); else G##j_dbl(&tab[i], &tab[i / 2]); } \
    G##j_set_inf(&acc);                                                                                         \
    for (int w = 63; w >= 0; w--) {                                                                             \
      G##j_dbl(&acc, &acc); G##j_dbl(&acc, &acc); G##j_dbl(&acc, &acc); G##j_dbl(&acc, &acc);                   \
      unsigned d = (unsigned)((k[w >> 4] >> ((w & 15) * 4)) & 15);                                              \
      if (d) G##j_add(&acc, &acc, &tab[d]);                                                                     \
    }                                                                                                           \
    G##j_to_affine(r, &acc);                                                                                    \
  }                                                                                                             \
  /* Affine + Affine -> normalised affine (data_structures.rs:185-190) */                                      \
  static void G##a_add(G##a* r, const G##a* p, const G##a* q) {                                                 \
    G##j a, b; G##j_from_affine(&a, p); G##j_from_affine(&b, q); G##j_add(&a, &a, &b); G##j_to_affine(r, &a);   \
  }

CURVE_IMPL(g1, fp, fp, r->z = FP_ONE)
CURVE_IMPL(g2, fp2, fp2, r->z.c0 = FP_ONE)

/* Montgomery Fr -> canonical integer */
static void fr_canon(uint64_t out[4], const fr* a) {
  uint64_t t[5] = {a->l[0], a->l[1], a->l[2], a->l[3], 0};
  for (int i = 0; i < 4; i++) { /* Montgomery reduction by R: multiply by 1 */
    uint64_t m = t[0] * RINV;
    u128 c = (u128)m * RMOD[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * RMOD[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = (uint64_t)(c >> 64);
  }
  int ge = 1;
  for (int i = 3; i >= 0; i--) {
    if (t[i] > RMOD[i]) break;
    if (t[i] < RMOD[i]) { ge = 0; break; }
  }
  if (ge) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)t[i] - RMOD[i] - b;
      t[i] = (uint64_t)d;
      b = (d >> 64) & 1;
    }
  }
  memcpy(out, t, 32);
}

/* ------------------------------------------------------------------ pairing (arkworks-shaped) */
typedef struct { fp2 c0, c1, c2; } ell;
static void fp_half(fp* r, const fp* a) { /* a/2 mod p on the residue itself (valid in Montgomery form too) */
  uint64_t t[7];
  memcpy(t, a->l, 48);
  t[6] = 0;
  if (t[0] & 1) {
    u128 c = 0;
    for (int i = 0; i < 6; i++) {
      c += (u128)t[i] + P[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
    t[6] = (uint64_t)c;
  }
  for (int i = 0; i < 6; i++) r->l[i] = (t[i] >> 1) | (t[i + 1] << 63);
}
static void fp2_half(fp2* r, const fp2* a) { fp_half(&r->c0, &a->c0); fp_half(&r->c1, &a->c1); }
static void g2_prepare(ell* out, const g2a* q) {
  fp2 x = q->x, y = q->y, z;
  memset(&z, 0, sizeof z);
  z.c0 = FP_ONE;
  int idx = 0;
  for (int b = 62; b >= 0; b--) {
    fp2 a, bb, c, e, f, g, h, i, j, e2, s;
    fp2_mul(&a, &x, &y); fp2_half(&a, &a);
    fp2_sqr(&bb, &y); fp2_sqr(&c, &z);
    fp2_dbl(&s, &c); fp2_add(&s, &s, &c); fp2_mul_xi(&e, &s); fp2_dbl(&e, &e); fp2_dbl(&e, &e);
    fp2_dbl(&f, &e); fp2_add(&f, &f, &e);
    fp2_add(&g, &bb, &f); fp2_half(&g, &g);
    fp2_add(&h, &y, &z); fp2_sqr(&h, &h); fp2_add(&s, &bb, &c); fp2_sub(&h, &h, &s);
    fp2_sub(&i, &e, &bb); fp2_sqr(&j, &x); fp2_sqr(&e2, &e);
    fp2_sub(&s, &bb, &f); fp2_mul(&x, &a, &s);
    fp2_sqr(&g, &g); fp2_dbl(&s, &e2); fp2_add(&s, &s, &e2); fp2_sub(&y, &g, &s);
    fp2_mul(&z, &bb, &h);
    out[idx].c0 = i; fp2_dbl(&s, &j); fp2_add(&out[idx].c1, &s, &j); fp2_neg(&out[idx].c2, &h);
    idx++;
    if ((X_ABS >> b) & 1) {
      fp2 th, la, cc, d, ee, ff, gg, hh, jj;
      fp2_mul(&s, &q->y, &z); fp2_sub(&th, &y, &s);
      fp2_mul(&s, &q->x, &z); fp2_sub(&la, &x, &s);
      fp2_sqr(&cc, &th); fp2_sqr(&d, &la); fp2_mul(&ee, &la, &d); fp2_mul(&ff, &z, &cc); fp2_mul(&gg, &x, &d);
      fp2_add(&hh, &ee, &ff); fp2_sub(&hh, &hh, &gg); fp2_sub(&hh, &hh, &gg);
      fp2_mul(&x, &la, &hh);
      fp2_sub(&s, &gg, &hh); fp2_mul(&s, &th, &s); fp2_mul(&jj, &ee, &y); fp2_sub(&y, &s, &jj);
      fp2_mul(&z, &z, &ee);
      fp2_mul(&jj, &th, &q->x); fp2_mul(&s, &la, &q->y); fp2_sub(&out[idx].c0, &jj, &s);
      fp2_neg(&out[idx].c1, &th); out[idx].c2 = la;
      idx++;
    }
  }
}
static void fp6_mul_by_01(fp6* r, const fp6* a, const fp2* b0, const fp2* b1) {
  fp2 aa, bb, t1, t2, t3, s;
  fp2_mul(&aa, &a->c0, b0); fp2_mul(&bb, &a->c1, b1);
  fp2_add(&s, &a->c1, &a->c2); fp2_mul(&t1, &s, b1); fp2_sub(&t1, &t1, &bb); fp2_mul_xi(&t1, &t1); fp2_add(&t1, &t1, &aa);
  fp2_add(&s, &a->c0, &a->c2); fp2_mul(&t3, &s, b0); fp2_sub(&t3, &t3, &aa); fp2_add(&t3, &t3, &bb);
  fp2_add(&t2, b0, b1); fp2_add(&s, &a->c0, &a->c1); fp2_mul(&t2, &t2, &s); fp2_sub(&t2, &t2, &aa); fp2_sub(&t2, &t2, &bb);
  r->c0 = t1; r->c1 = t2; r->c2 = t3;
}
static void fp6_mul_by_1(fp6* r, const fp6* a, const fp2* b1) {
  fp2 t0, t1, t2;
  fp2_mul(&t0, &a->c2, b1); fp2_mul_xi(&t0, &t0); fp2_mul(&t1, &a->c0, b1); fp2_mul(&t2, &a->c1, b1);
  r->c0 = t0; r->c1 = t1; r->c2 = t2;
}
static void mul_by_014(fp12* f, const fp2* c0, const fp2* c1, const fp2* c4) { /* 13 Fp2 products, as ark-ff */
  fp6 aa, bb, s;
  fp2 o;
  fp6_mul_by_01(&aa, &f->c0, c0, c1);
  fp6_mul_by_1(&bb, &f->c1, c4);
  fp2_add(&o, c1, c4);
  fp6_add(&s, &f->c0, &f->c1);
  fp6_mul_by_01(&s, &s, c0, &o);
  fp6_sub(&s, &s, &aa);
  fp6_sub(&f->c1, &s, &bb);
  fp6_mul_v(&bb, &bb);
  fp6_add(&f->c0, &aa, &bb);
}
/* multi_miller_loop over n pairs (identity pairs dropped) */
static void multi_miller(fp12* f, int n, const g1a* ps, const g2a* qs) {
  ell* lines = (ell*)malloc((size_t)(n ? n : 1) * 68 * sizeof(ell));
  const g1a** pp = (const g1a**)malloc((size_t)(n ? n : 1) * sizeof(void*));
  int cnt = 0;
  for (int i = 0; i < n; i++) {
    if (g1a_inf(&ps[i]) || g2a_inf(&qs[i])) continue;
    g2_prepare(lines + (size_t)cnt * 68, &qs[i]);
    pp[cnt++] = &ps[i];
  }
  fp12_one(f);
  int idx = 0;
  for (int b = 62; b >= 0; b--) {
    fp12_sqr(f, f);
    int nl = ((X_ABS >> b) & 1) ? 2 : 1;
    for (int t = 0; t < nl; t++, idx++)
      for (int k = 0; k < cnt; k++) {
        const ell* l = &lines[(size_t)k * 68 + idx];
        fp2 c1, c2;
        fp2_mul_fp(&c1, &l->c1, &pp[k]->x);
        fp2_mul_fp(&c2, &l->c2, &pp[k]->y);
        mul_by_014(f, &l->c0, &c1, &c2);
      }
  }
  fp12_conj(f, f);
  free(lines);
  free(pp);
}
static void exp_x(fp12* r, const fp12* a) { /* a^x, x < 0, a in the cyclotomic subgroup */
  fp12 acc = *a;
  for (int b = 62; b >= 0; b--) {
    fp12_cyclo_sqr(&acc, &acc);
    if ((X_ABS >> b) & 1) fp12_mul(&acc, &acc, a);
  }
  fp12_conj(r, &acc);
}
static void final_exp(fp12* out, const fp12* f) {
  fp12 r, t, a, b, c;
  fp12_inv(&t, f); fp12_conj(&r, f); fp12_mul(&r, &r, &t);
  fp12_frob(&t, &r, 2); fp12_mul(&r, &r, &t);
  exp_x(&a, &r); fp12_conj(&t, &r); fp12_mul(&a, &a, &t);
  exp_x(&b, &a); fp12_conj(&t, &a); fp12_mul(&b, &b, &t);
  exp_x(&c, &b); fp12_frob(&t, &b, 1); fp12_mul(&c, &c, &t);
  exp_x(&a, &c); exp_x(&a, &a); fp12_frob(&t, &c, 2); fp12_mul(&a, &a, &t); fp12_conj(&t, &c); fp12_mul(&a, &a, &t);
  fp12_cyclo_sqr(&t, &r); fp12_mul(&t, &t, &r);
  fp12_mul(out, &a, &t);
}
static void multi_pairing(fp12* out, int n, const g1a* ps, const g2a* qs) {
  fp12 f;
  multi_miller(&f, n, ps, qs);
  final_exp(out, &f);
}

/* ------------------------------------------------------------------ GS layer */
typedef struct { g1a p[2]; } com1;
typedef struct { g2a p[2]; } com2;
typedef struct { fp12 e[4]; } comt;
typedef struct { com1 u[2]; com2 v[2]; g1a g1; g2a g2; fp12 gt; } crs_t;

static void com1_smul(com1* r, const com1* a, const uint64_t k[4]) { g1a_mul(&r->p[0], &a->p[0], k); g1a_mul(&r->p[1], &a->p[1], k); }
static void com2_smul(com2* r, const com2* a, const uint64_t k[4]) { g2a_mul(&r->p[0], &a->p[0], k); g2a_mul(&r->p[1], &a->p[1], k); }
static void com1_add(com1* r, const com1* a, const com1* b) { g1a_add(&r->p[0], &a->p[0], &b->p[0]); g1a_add(&r->p[1], &a->p[1], &b->p[1]); }
static void com2_add(com2* r, const com2* a, const com2* b) { g2a_add(&r->p[0], &a->p[0], &b->p[0]); g2a_add(&r->p[1], &a->p[1], &b->p[1]); }

/* ComT::pairing_sum: 4 multi_pairings (data_structures.rs:494-502) */
static void pairing_sum(comt* out, int k, const com1* xs, const com2* ys) {
  g1a* ps = (g1a*)malloc((size_t)(k ? k : 1) * sizeof(g1a));
  g2a* qs = (g2a*)malloc((size_t)(k ? k : 1) * sizeof(g2a));
  for (int a = 0; a < 2; a++)
    for (int b = 0; b < 2; b++) {
      for (int i = 0; i < k; i++) { ps[i] = xs[i].p[a]; qs[i] = ys[i].p[b]; }
      multi_pairing(&out->e[2 * a + b], k, ps, qs);
    }
  free(ps);
  free(qs);
}
static void comt_mul(comt* r, const comt* a, const comt* b) { for (int i = 0; i < 4; i++) fp12_mul(&r->e[i], &a->e[i], &b->e[i]); }

/* PPE::verify (verifier.rs:23-55), reference evaluation order */
static int verify_ppe(int m, int n, const g1a* A, const g2a* B, const fr* gamma, const fp12* target, const com1* c,
                      const com2* d, const com2* pi, const com1* theta, const crs_t* crs) {
  com1* linA = (com1*)calloc((size_t)n, sizeof(com1));
  com2* linB = (com2*)calloc((size_t)m, sizeof(com2));
  com2* gd = (com2*)calloc((size_t)m, sizeof(com2));
  for (int j = 0; j < n; j++) linA[j].p[1] = A[j];
  for (int i = 0; i < m; i++) linB[i].p[1] = B[i];
  comt t1, t2, t3, t4, t5, lhs, rhs;
  pairing_sum(&t1, n, linA, d);
  pairing_sum(&t2, m, c, linB);
  for (int i = 0; i < m; i++) { /* Gamma * d: m*n Com2 scalar muls, summed term by term (:39-40) */
    com2 acc;
    memset(&acc, 0, sizeof acc);
    for (int j = 0; j < n; j++) {
      uint64_t k[4];
      com2 t;
      fr_canon(k, &gamma[i * n + j]);
      com2_smul(&t, &d[j], k);
      com2_add(&acc, &acc, &t);
    }
    gd[i] = acc;
  }
  pairing_sum(&t3, m, c, gd);
  pairing_sum(&t4, 2, crs->u, pi);
  pairing_sum(&t5, 2, theta, crs->v);
  comt_mul(&lhs, &t1, &t2);
  comt_mul(&lhs, &lhs, &t3);
  comt lin_t;
  for (int i = 0; i < 3; i++) fp12_one(&lin_t.e[i]);
  lin_t.e[3] = *target;
  comt_mul(&rhs, &lin_t, &t4);
  comt_mul(&rhs, &rhs, &t5);
  free(linA);
  free(linB);
  free(gd);
  for (int i = 0; i < 4; i++)
    if (!fp12_eq(&lhs.e[i], &rhs.e[i])) return 0;
  return 1;
}

/* ------------------------------------------------------------------ exported API (ctypes) */
void gsref_pairing(const void* p, const void* q, void* out) { multi_pairing((fp12*)out, 1, (const g1a*)p, (const g2a*)q); }
void gsref_pairing_sum(int k, const void* xs, const void* ys, void* out) { pairing_sum((comt*)out, k, (const com1*)xs, (const com2*)ys); }
void gsref_g1_mul(const void* p, const void* k_mont, void* out) { uint64_t k[4]; fr_canon(k, (const fr*)k_mont); g1a_mul((g1a*)out, (const g1a*)p, k); }
void gsref_g2_mul(const void* p, const void* k_mont, void* out) { uint64_t k[4]; fr_canon(k, (const fr*)k_mont); g2a_mul((g2a*)out, (const g2a*)p, k); }

/* batch_commit_G1 / G2 (commit.rs:78-100, 178-200): c_i = (O, X_i) + R[i][0] u1 + R[i][1] u2, term by term */
void gsref_batch_commit_g1(size_t n, const void* xvars, const void* rand, const void* crs_, void* out) {
  const crs_t* crs = (const crs_t*)crs_;
  const g1a* X = (const g1a*)xvars;
  const fr* R = (const fr*)rand;
  com1* o = (com1*)out;
  for (size_t i = 0; i < n; i++) {
    com1 acc, t, lin;
    memset(&acc, 0, sizeof acc);
    for (int k = 0; k < 2; k++) {
      uint64_t s[4];
      fr_canon(s, &R[2 * i + k]);
      com1_smul(&t, &crs->u[k], s);
      com1_add(&acc, &acc, &t);
    }
    memset(&lin, 0, sizeof lin);
    lin.p[1] = X[i];
    com1_add(&o[i], &lin, &acc);
  }
}
void gsref_batch_commit_g2(size_t n, const void* yvars, const void* rand, const void* crs_, void* out) {
  const crs_t* crs = (const crs_t*)crs_;
  const g2a* Y = (const g2a*)yvars;
  const fr* S = (const fr*)rand;
  com2* o = (com2*)out;
  for (size_t i = 0; i < n; i++) {
    com2 acc, t, lin;
    memset(&acc, 0, sizeof acc);
    for (int k = 0; k < 2; k++) {
      uint64_t s[4];
      fr_canon(s, &S[2 * i + k]);
      com2_smul(&t, &crs->v[k], s);
      com2_add(&acc, &acc, &t);
    }
    memset(&lin, 0, sizeof lin);
    lin.p[1] = Y[i];
    com2_add(&o[i], &lin, &acc);
  }
}

typedef struct {
  int m, n;
  size_t lo, hi;
  const g1a* A; const g2a* B; const fr* gamma; const fp12* target;
  const com1* c; const com2* d; const com2* pi; const com1* theta;
  const crs_t* crs;
  uint8_t* ok;
} vjob;
static void* vworker(void* arg) {
  vjob* j = (vjob*)arg;
  for (size_t p = j->lo; p < j->hi; p++)
    j->ok[p] = (uint8_t)verify_ppe(j->m, j->n, j->A + p * j->n, j->B + p * j->m, j->gamma + p * j->m * j->n, j->target + p,
                                   j->c + p * j->m, j->d + p * j->n, j->pi + p * 2, j->theta + p * 2, j->crs);
  return 0;
}
/* `count` independent PPE verifications spread over `nthreads` host threads (same array layout as gs_verify_batch) */
void gsref_verify_ppe_batch(size_t count, int m, int n, const void* A, const void* B, const void* gamma, const void* target,
                            const void* c, const void* d, const void* pi, const void* theta, const void* crs, uint8_t* ok,
                            int nthreads) {
  if (nthreads < 1) nthreads = 1;
  pthread_once(&frob_once, frob_init);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  vjob* jobs = (vjob*)malloc(sizeof(vjob) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    vjob j = {m, n, count * t / nthreads, count * (t + 1) / nthreads, (const g1a*)A, (const g2a*)B, (const fr*)gamma,
              (const fp12*)target, (const com1*)c, (const com2*)d, (const com2*)pi, (const com1*)theta, (const crs_t*)crs, ok};
    jobs[t] = j;
    pthread_create(&th[t], 0, vworker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
  free(th);
  free(jobs);
}

/* ================================================================== round 2: the whole prove / verify path
 * Everything below restates the reference with its own evaluation structure; `nthreads` only spreads the
 * independent Com::scalar_mul calls of one left_mul over host threads. */

/* ------------------------------------------------------------------ Fr (Montgomery, 4 x 64, R = 2^256) */
static void fr_mulm(fr* r, const fr* a, const fr* b) { /* CIOS */
  uint64_t t[6] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * RINV;
    c = (u128)m * RMOD[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * RMOD[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  int ge = t[4] != 0;
  if (!ge) {
    ge = 1;
    for (int i = 3; i >= 0; i--) {
      if (t[i] > RMOD[i]) break;
      if (t[i] < RMOD[i]) { ge = 0; break; }
    }
  }
  if (ge) {
    u128 b2 = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)t[i] - RMOD[i] - b2;
      t[i] = (uint64_t)d;
      b2 = (d >> 64) & 1;
    }
  }
  memcpy(r->l, t, 32);
}
static void fr_addm(fr* r, const fr* a, const fr* b) {
  u128 c = 0;
  uint64_t t[4];
  for (int i = 0; i < 4; i++) {
    c += (u128)a->l[i] + b->l[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  int ge = 1; /* r < 2^255, no carry out */
  for (int i = 3; i >= 0; i--) {
    if (t[i] > RMOD[i]) break;
    if (t[i] < RMOD[i]) { ge = 0; break; }
  }
  if (ge) {
    u128 b2 = 0;
    for (int i = 0; i < 4; i++) {
      u128 d = (u128)t[i] - RMOD[i] - b2;
      t[i] = (uint64_t)d;
      b2 = (d >> 64) & 1;
    }
  }
  memcpy(r->l, t, 32);
}
static void fr_negm(fr* r, const fr* a) {
  if ((a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0) { *r = *a; return; }
  u128 b2 = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)RMOD[i] - a->l[i] - b2;
    r->l[i] = (uint64_t)d;
    b2 = (d >> 64) & 1;
  }
}
/* Matrix<Fr>::right_mul (data_structures.rs:824-868): out (r x c) = a (r x k) * b (k x c), row-major */
static void frmat_mul(fr* out, const fr* a, const fr* b, size_t r, size_t k, size_t c) {
  for (size_t i = 0; i < r; i++)
    for (size_t j = 0; j < c; j++) {
      fr acc, t;
      memset(&acc, 0, sizeof acc);
      for (size_t l = 0; l < k; l++) {
        fr_mulm(&t, &a[i * k + l], &b[l * c + j]);
        fr_addm(&acc, &acc, &t);
      }
      out[i * c + j] = acc;
    }
}
static void frmat_transpose(fr* out, const fr* a, size_t r, size_t c) {
  for (size_t i = 0; i < r; i++)
    for (size_t j = 0; j < c; j++) out[j * r + i] = a[i * c + j];
}

/* ------------------------------------------------------------------ host-thread helper */
typedef void (*pf_fn)(size_t i, void* arg);
typedef struct { size_t n; size_t next; pf_fn fn; void* arg; pthread_mutex_t mu; } pf_state;
static void* pf_worker(void* p) {
  pf_state* s = (pf_state*)p;
  for (;;) {
    pthread_mutex_lock(&s->mu);
    size_t i = s->next++;
    pthread_mutex_unlock(&s->mu);
    if (i >= s->n) return 0;
    s->fn(i, s->arg);
  }
}
static void parallel_for(size_t n, int nthreads, pf_fn fn, void* arg) {
  pthread_once(&frob_once, frob_init);
  pthread_once(&r2_once, r2_init);
  if (nthreads <= 1 || n <= 1) {
    for (size_t i = 0; i < n; i++) fn(i, arg);
    return;
  }
  pf_state s = {n, 0, fn, arg, PTHREAD_MUTEX_INITIALIZER};
  if ((size_t)nthreads > n) nthreads = (int)n;
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 0; t < nthreads; t++) pthread_create(&th[t], 0, pf_worker, &s);
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
  free(th);
}

/* ------------------------------------------------------------------ Mat::left_mul on Matrix<Com> (:696-742)
 * mat is a column vector of k Com, lhs is rows x k (row-major): out[i] = sum_k mat[k].scalar_mul(lhs[i][k]),
 * the sum folded from Com::zero() with one affine add (+ normalisation) per term (:244-250, :185-190). */
#define LEFT_MUL_IMPL(C)                                                                                       \
  typedef struct { const C* mat; const fr* lhs; C* terms; size_t k; } C##_lm_job;                              \
  static void C##_lm_term(size_t idx, void* arg) {                                                             \
    C##_lm_job* j = (C##_lm_job*)arg;                                                                          \
    uint64_t s[4];                                                                                             \
    fr_canon(s, &j->lhs[idx]);                                                                                 \
    C##_smul(&j->terms[idx], &j->mat[idx % j->k], s);                                                          \
  }                                                                                                            \
  static void C##_left_mul(C* out, const C* mat, const fr* lhs, size_t rows, size_t k, int nthreads) {         \
    C* terms = (C*)malloc(sizeof(C) * (rows * k + 1));                                              \
    C##_lm_job job = {mat, lhs, terms, k};                                                                     \
    parallel_for(rows * k, nthreads, C##_lm_term, &job);                                                       \
    for (size_t i = 0; i < rows; i++) {                                                                        \
      C acc;                                                                                                   \
      memset(&acc, 0, sizeof acc);                                                                             \
      for (size_t l = 0; l < k; l++) C##_add(&acc, &acc, &terms[i * k + l]);                                   \
      out[i] = acc;                                                                                            \
    }                                                                                                          \
    free(terms);                                                                                               \
  }
LEFT_MUL_IMPL(com1)
LEFT_MUL_IMPL(com2)

/* W1 = u2 + iota_1(g1), W2 = v2 + iota_2(g2) (data_structures.rs:325, :370) */
static void crs_w1(com1* w, const crs_t* crs) { com1 lin; memset(&lin, 0, sizeof lin); lin.p[1] = crs->g1; com1_add(w, &crs->u[1], &lin); }
static void crs_w2(com2* w, const crs_t* crs) { com2 lin; memset(&lin, 0, sizeof lin); lin.p[1] = crs->g2; com2_add(w, &crs->v[1], &lin); }

/* batch_linear_map (:315-320, :360-365) or batch_scalar_linear_map (:328-334, :373-379), by side kind */
typedef struct { const fr* s; const com1* w1; const com2* w2; com1* o1; com2* o2; } slm_job;
static void slm1_term(size_t i, void* arg) { slm_job* j = (slm_job*)arg; uint64_t k[4]; fr_canon(k, &j->s[i]); com1_smul(&j->o1[i], j->w1, k); }
static void slm2_term(size_t i, void* arg) { slm_job* j = (slm_job*)arg; uint64_t k[4]; fr_canon(k, &j->s[i]); com2_smul(&j->o2[i], j->w2, k); }
static com1* map_side1(int is_group, size_t cnt, const void* elems, const crs_t* crs, int nthreads) {
  com1* out = (com1*)calloc(cnt ? cnt : 1, sizeof(com1));
  if (is_group) { for (size_t i = 0; i < cnt; i++) out[i].p[1] = ((const g1a*)elems)[i]; return out; }
  com1 w; crs_w1(&w, crs);
  slm_job j = {(const fr*)elems, &w, 0, out, 0};
  parallel_for(cnt, nthreads, slm1_term, &j);
  return out;
}
static com2* map_side2(int is_group, size_t cnt, const void* elems, const crs_t* crs, int nthreads) {
  com2* out = (com2*)calloc(cnt ? cnt : 1, sizeof(com2));
  if (is_group) { for (size_t i = 0; i < cnt; i++) out[i].p[1] = ((const g2a*)elems)[i]; return out; }
  com2 w; crs_w2(&w, crs);
  slm_job j = {(const fr*)elems, 0, &w, 0, out};
  parallel_for(cnt, nthreads, slm2_term, &j);
  return out;
}

/* batch_commit_scalar_to_B1 / B2 (commit.rs:125-156, 225-256): c_i = iota'(x_i) + r_i u1 */
void gsref_batch_commit_scalar_b1(size_t n, const void* xs, const void* rand, const void* crs_, void* out, int nthreads) {
  const crs_t* crs = (const crs_t*)crs_;
  com1* lin = map_side1(0, n, xs, crs, nthreads);
  com1* ru = (com1*)malloc(sizeof(com1) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) com1_left_mul(&ru[i], &crs->u[0], (const fr*)rand + i, 1, 1, 1);
  for (size_t i = 0; i < n; i++) com1_add(&((com1*)out)[i], &lin[i], &ru[i]);
  free(lin);
  free(ru);
}
void gsref_batch_commit_scalar_b2(size_t n, const void* ys, const void* rand, const void* crs_, void* out, int nthreads) {
  const crs_t* crs = (const crs_t*)crs_;
  com2* lin = map_side2(0, n, ys, crs, nthreads);
  com2* rv = (com2*)malloc(sizeof(com2) * (n ? n : 1));
  for (size_t i = 0; i < n; i++) com2_left_mul(&rv[i], &crs->v[0], (const fr*)rand + i, 1, 1, 1);
  for (size_t i = 0; i < n; i++) com2_add(&((com2*)out)[i], &lin[i], &rv[i]);
  free(lin);
  free(rv);
}

/* Provable::prove (prove.rs:92-171 PPE, 195-274 MSMEG1, 298-379 MSMEG2, 409-488 Quad).
 * type = EquType byte; x side is G1 points for types 0,1 and Fr for 2,3; y side is G2 points for 0,2, Fr for 1,3.
 * x_rand m x cx, y_rand n x cy, pf_rand = T (cy x cx row-major, the draw order).  out_pi: cx Com2, out_theta: cy Com1. */
void gsref_prove(int type, size_t m, size_t n, const void* a_consts, const void* b_consts, const void* gamma_,
                 const void* xvars, const void* yvars, const void* x_rand, const void* y_rand, const void* pf_rand,
                 const void* crs_, void* out_pi, void* out_theta, int nthreads) {
  const crs_t* crs = (const crs_t*)crs_;
  const fr* gamma = (const fr*)gamma_;
  const fr* T = (const fr*)pf_rand;
  int xg = (type == 0 || type == 1), yg = (type == 0 || type == 2);
  size_t cx = xg ? 2 : 1, cy = yg ? 2 : 1;
  fr* Rt = (fr*)malloc(sizeof(fr) * cx * m);
  fr* St = (fr*)malloc(sizeof(fr) * cy * n);
  frmat_transpose(Rt, (const fr*)x_rand, m, cx); /* x_rand_trans: cx x m */
  frmat_transpose(St, (const fr*)y_rand, n, cy); /* y_rand_trans: cy x n */

  /* ---- pi */
  com2 lin_b_part[2], stmt_y_part[2], key_part[2];
  com2* linB = map_side2(yg, m, b_consts, crs, nthreads);
  com2_left_mul(lin_b_part, linB, Rt, cx, m, nthreads);              /* x_rand_lin_b */
  fr* RtG = (fr*)malloc(sizeof(fr) * cx * n);
  frmat_mul(RtG, Rt, gamma, cx, m, n);                               /* x_rand_stmt = R^T Gamma */
  com2* linY = map_side2(yg, n, yvars, crs, nthreads);
  com2_left_mul(stmt_y_part, linY, RtG, cx, n, nthreads);            /* x_rand_stmt_lin_y */
  fr RtGS[4], Tt[4];
  frmat_mul(RtGS, RtG, (const fr*)y_rand, cx, n, cy);                /* (R^T Gamma) S : cx x cy */
  frmat_transpose(Tt, T, cy, cx);                                    /* T^T : cx x cy */
  for (size_t i = 0; i < cx * cy; i++) { fr t; fr_negm(&t, &Tt[i]); fr_addm(&RtGS[i], &RtGS[i], &t); }
  com2_left_mul(key_part, crs->v, RtGS, cx, cy, nthreads);           /* v (or [v1]) . pf_rand_stmt */
  for (size_t i = 0; i < cx; i++) {
    com2 t;
    com2_add(&t, &lin_b_part[i], &stmt_y_part[i]);
    com2_add(&((com2*)out_pi)[i], &t, &key_part[i]);
  }

  /* ---- theta */
  com1 lin_a_part[2], stmt_x_part[2], ukey_part[2];
  com1* linA = map_side1(xg, n, a_consts, crs, nthreads);
  com1_left_mul(lin_a_part, linA, St, cy, n, nthreads);              /* y_rand_lin_a */
  fr* Gt = (fr*)malloc(sizeof(fr) * (m * n + 1));
  frmat_transpose(Gt, gamma, m, n);
  fr* StGt = (fr*)malloc(sizeof(fr) * cy * m);
  frmat_mul(StGt, St, Gt, cy, n, m);                                 /* y_rand_stmt = S^T Gamma^T */
  com1* linX = map_side1(xg, m, xvars, crs, nthreads);
  com1_left_mul(stmt_x_part, linX, StGt, cy, m, nthreads);           /* y_rand_stmt_lin_x */
  com1_left_mul(ukey_part, crs->u, T, cy, cx, nthreads);             /* u (or [u1]) . pf_rand */
  for (size_t i = 0; i < cy; i++) {
    com1 t;
    com1_add(&t, &lin_a_part[i], &stmt_x_part[i]);
    com1_add(&((com1*)out_theta)[i], &t, &ukey_part[i]);
  }
  free(Rt); free(St); free(linB); free(RtG); free(linY); free(linA); free(Gt); free(StGt); free(linX);
}

/* ComT::pairing (:484-491): four separate full pairings */
static void comt_pairing(comt* out, const com1* x, const com2* y) {
  for (int a = 0; a < 2; a++)
    for (int b = 0; b < 2; b++) multi_pairing(&out->e[2 * a + b], 1, &x->p[a], &y->p[b]);
}
/* pairing_sum with its four multi_pairings on separate host threads (the reference runs them one after the other) */
typedef struct { comt* out; int k; const com1* xs; const com2* ys; } ps_job;
static void ps_entry(size_t e, void* arg) {
  ps_job* j = (ps_job*)arg;
  int a = (int)e >> 1, b = (int)e & 1;
  g1a* ps = (g1a*)malloc((size_t)(j->k ? j->k : 1) * sizeof(g1a));
  g2a* qs = (g2a*)malloc((size_t)(j->k ? j->k : 1) * sizeof(g2a));
  for (int i = 0; i < j->k; i++) { ps[i] = j->xs[i].p[a]; qs[i] = j->ys[i].p[b]; }
  multi_pairing(&j->out->e[e], j->k, ps, qs);
  free(ps);
  free(qs);
}
static void pairing_sum_mt(comt* out, int k, const com1* xs, const com2* ys, int nthreads) {
  ps_job j = {out, k, xs, ys};
  parallel_for(4, nthreads, ps_entry, &j);
}

/* Verifiable::verify (verifier.rs:23-55, 57-89, 91-123, 125-157) for one (equation, proof). */
static int verify_any(int type, size_t m, size_t n, const void* a_consts, const void* b_consts, const fr* gamma,
                      const void* target, const com1* c, const com2* d, const com2* pi, const com1* theta,
                      const crs_t* crs, int nthreads) {
  int xg = (type == 0 || type == 1), yg = (type == 0 || type == 2);
  com1* linA = map_side1(xg, n, a_consts, crs, nthreads);
  com2* linB = map_side2(yg, m, b_consts, crs, nthreads);
  comt t1, t2, t3, t4, t5, lin_t, lhs, rhs;
  pairing_sum_mt(&t1, (int)n, linA, d, nthreads);                    /* lin_a_com_y */
  pairing_sum_mt(&t2, (int)m, c, linB, nthreads);                    /* com_x_lin_b */
  com2* gd = (com2*)malloc(sizeof(com2) * (m ? m : 1));
  com2_left_mul(gd, d, gamma, m, n, nthreads);                       /* stmt_com_y = Gamma . d (:39-40) */
  pairing_sum_mt(&t3, (int)m, c, gd, nthreads);                      /* com_x_stmt_com_y */
  com1 w1; com2 w2;
  crs_w1(&w1, crs);
  crs_w2(&w2, crs);
  if (type == 0) {                                                   /* linear_map_PPE :509-516 */
    for (int i = 0; i < 3; i++) fp12_one(&lin_t.e[i]);
    lin_t.e[3] = *(const fp12*)target;
  } else if (type == 1) {                                            /* linear_map_MSMEG1 :519-524 */
    com1 lt; memset(&lt, 0, sizeof lt); lt.p[1] = *(const g1a*)target;
    comt_pairing(&lin_t, &lt, &w2);
  } else if (type == 2) {                                            /* linear_map_MSMEG2 :527-532 */
    com2 lt; memset(&lt, 0, sizeof lt); lt.p[1] = *(const g2a*)target;
    comt_pairing(&lin_t, &w1, &lt);
  } else {                                                           /* linear_map_quad :535-540 */
    uint64_t k[4]; com2 tw;
    fr_canon(k, (const fr*)target);
    com2_smul(&tw, &w2, k);
    comt_pairing(&lin_t, &w1, &tw);
  }
  if (xg) pairing_sum_mt(&t4, 2, crs->u, pi, nthreads); else comt_pairing(&t4, &crs->u[0], &pi[0]);
  if (yg) pairing_sum_mt(&t5, 2, theta, crs->v, nthreads); else comt_pairing(&t5, &theta[0], &crs->v[0]);
  comt_mul(&lhs, &t1, &t2);
  comt_mul(&lhs, &lhs, &t3);
  comt_mul(&rhs, &lin_t, &t4);
  comt_mul(&rhs, &rhs, &t5);
  free(linA); free(linB); free(gd);
  for (int i = 0; i < 4; i++)
    if (!fp12_eq(&lhs.e[i], &rhs.e[i])) return 0;
  return 1;
}
int gsref_verify(int type, size_t m, size_t n, const void* a_consts, const void* b_consts, const void* gamma,
                 const void* target, const void* xcoms, const void* ycoms, const void* pi, const void* theta,
                 const void* crs, int nthreads) {
  pthread_once(&frob_once, frob_init);
  pthread_once(&r2_once, r2_init);
  return verify_any(type, m, n, a_consts, b_consts, (const fr*)gamma, target, (const com1*)xcoms, (const com2*)ycoms,
                    (const com2*)pi, (const com1*)theta, (const crs_t*)crs, nthreads);
}

/* `count` independent verifications of one type and shape over `nthreads` host threads (array layout of
 * gs_verify_batch; each verification itself single-threaded) */
typedef struct {
  int type; size_t m, n;
  const uint8_t *A, *B, *gamma, *target, *c, *d, *pi, *theta;
  size_t sa, sb, st, cx, cy;
  const crs_t* crs; uint8_t* ok;
} vbjob;
static void vb_one(size_t p, void* arg) {
  vbjob* j = (vbjob*)arg;
  j->ok[p] = (uint8_t)verify_any(j->type, j->m, j->n, j->A + p * j->n * j->sa, j->B + p * j->m * j->sb,
                                 (const fr*)(j->gamma + p * j->m * j->n * 32), j->target + p * j->st,
                                 (const com1*)(j->c + p * j->m * 192), (const com2*)(j->d + p * j->n * 384),
                                 (const com2*)(j->pi + p * j->cx * 384), (const com1*)(j->theta + p * j->cy * 192), j->crs, 1);
}
void gsref_verify_batch(int type, size_t count, size_t m, size_t n, const void* A, const void* B, const void* gamma,
                        const void* target, const void* c, const void* d, const void* pi, const void* theta, const void* crs,
                        uint8_t* ok, int nthreads) {
  int xg = (type == 0 || type == 1), yg = (type == 0 || type == 2);
  static const size_t tsz[4] = {576, 96, 192, 32};
  vbjob j = {type, m, n, (const uint8_t*)A, (const uint8_t*)B, (const uint8_t*)gamma, (const uint8_t*)target,
             (const uint8_t*)c, (const uint8_t*)d, (const uint8_t*)pi, (const uint8_t*)theta,
             xg ? 96u : 32u, yg ? 192u : 32u, tsz[type], xg ? 2u : 1u, yg ? 2u : 1u, (const crs_t*)crs, ok};
  parallel_for(count, nthreads, vb_one, &j);
}

/* `count` independent Provable::prove calls (array layout of gs_prove_batch; shared_vars as there), spread over
 * host threads proof by proof */
typedef struct {
  int type; size_t m, n; int shared;
  const uint8_t *A, *B, *gamma, *X, *Y, *R, *S, *T;
  size_t sx, sy, cx, cy;
  const crs_t* crs; uint8_t *pi, *theta;
} pbjob;
static void pb_one(size_t p, void* arg) {
  pbjob* j = (pbjob*)arg;
  size_t q = j->shared ? 0 : p;
  gsref_prove(j->type, j->m, j->n, j->A + p * j->n * j->sx, j->B + p * j->m * j->sy, j->gamma + p * j->m * j->n * 32,
              j->X + q * j->m * j->sx, j->Y + q * j->n * j->sy, j->R + q * j->m * j->cx * 32, j->S + q * j->n * j->cy * 32,
              j->T + p * j->cx * j->cy * 32, j->crs, j->pi + p * j->cx * 384, j->theta + p * j->cy * 192, 1);
}
void gsref_prove_batch(int type, size_t count, size_t m, size_t n, const void* A, const void* B, const void* gamma,
                       const void* X, const void* Y, const void* R, const void* S, const void* T, int shared_vars,
                       const void* crs, void* out_pi, void* out_theta, int nthreads) {
  int xg = (type == 0 || type == 1), yg = (type == 0 || type == 2);
  pbjob j = {type, m, n, shared_vars, (const uint8_t*)A, (const uint8_t*)B, (const uint8_t*)gamma, (const uint8_t*)X,
             (const uint8_t*)Y, (const uint8_t*)R, (const uint8_t*)S, (const uint8_t*)T, xg ? 96u : 32u, yg ? 192u : 32u,
             xg ? 2u : 1u, yg ? 2u : 1u, (const crs_t*)crs, (uint8_t*)out_pi, (uint8_t*)out_theta};
  parallel_for(count, nthreads, pb_one, &j);
}

/* n scalar multiplications of ONE base (input generation for the big parity cases: points = k . g) */
typedef struct { const void* base; const fr* k; uint8_t* out; int g2; } mb_job;
static void mb_one(size_t i, void* arg) {
  mb_job* j = (mb_job*)arg;
  uint64_t s[4];
  fr_canon(s, &j->k[i]);
  if (j->g2) g2a_mul((g2a*)(j->out + i * 192), (const g2a*)j->base, s);
  else g1a_mul((g1a*)(j->out + i * 96), (const g1a*)j->base, s);
}
void gsref_g1_mul_batch(size_t n, const void* base, const void* ks, void* out, int nthreads) {
  mb_job j = {base, (const fr*)ks, (uint8_t*)out, 0};
  parallel_for(n, nthreads, mb_one, &j);
}
void gsref_g2_mul_batch(size_t n, const void* base, const void* ks, void* out, int nthreads) {
  mb_job j = {base, (const fr*)ks, (uint8_t*)out, 1};
  parallel_for(n, nthreads, mb_one, &j);
}
/* cross-check hook for the two inversion algorithms */
int gsref_selftest_inv(const void* a) {
  fp x, y;
  pthread_once(&r2_once, r2_init);
  fp_inv(&x, (const fp*)a);
  fp_inv_fermat(&y, (const fp*)a);
  return fp_eq(&x, &y);
}
