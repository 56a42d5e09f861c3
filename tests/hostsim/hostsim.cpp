// Host instantiation of the DEVICE math headers (tests only; never linked into the product).
// Lets the CPU-only CI check the exact limb algorithms / tower / curve / pairing formulas
// that the CUDA kernels execute, against oracle/ (a separately written big-int restatement).
#include <cstring>
#include <vector>
#include <cfenv>
#include "../../groth-sahai-rs_b200/csrc/pairing.cuh"
#include "../../tools/experimental/fpd.cuh"
#include "../../groth-sahai-rs_b200/csrc/endo.cuh"
#include "../../groth-sahai-rs_b200/csrc/wire.cuh"
using namespace gs;

#define LD(T, v, p) T v; memcpy(&v, p, sizeof(T))
#define ST(p, v) memcpy(p, &v, sizeof(v))
extern "C" {
void hs_fp_mul(void* r, const void* a, const void* b) { LD(fp, x, a); LD(fp, y, b); fp z; fp::mul(z, x, y); ST(r, z); }
// mulsum: nt in {1,2,3,4,6,8}; a, b arrays of nt fp
void hs_fp_mulsum(void* r, int nt, const void* a, const void* b) {
  const fp* A = (const fp*)a; const fp* B = (const fp*)b; fp z;
#define MS(NT) case NT: { fp x[NT], y[NT]; for (int i = 0; i < NT; i++) { x[i] = A[i]; y[i] = B[i]; } fp::mulsum<NT>(z, x, y); } break;
  switch (nt) { MS(1) MS(2) MS(3) MS(4) MS(6) MS(8) default: z.set_zero(); }
#undef MS
  ST(r, z); }
void hs_fp_add(void* r, const void* a, const void* b) { LD(fp, x, a); LD(fp, y, b); fp z; fp::add(z, x, y); ST(r, z); }
void hs_fp_sub(void* r, const void* a, const void* b) { LD(fp, x, a); LD(fp, y, b); fp z; fp::sub(z, x, y); ST(r, z); }
void hs_fp_neg(void* r, const void* a) { LD(fp, x, a); fp z; fp::neg(z, x); ST(r, z); }
void hs_fp_inv(void* r, const void* a) { LD(fp, x, a); fp z; fp_inv(z, x); ST(r, z); }
void hs_fp_inv_fermat(void* r, const void* a) { LD(fp, x, a); fp z; fp_inv_fermat(z, x); ST(r, z); }
void hs_fr_mul(void* r, const void* a, const void* b) { LD(fr, x, a); LD(fr, y, b); fr z; fr::mul(z, x, y); ST(r, z); }
void hs_fr_add(void* r, const void* a, const void* b) { LD(fr, x, a); LD(fr, y, b); fr z; fr::add(z, x, y); ST(r, z); }
void hs_fr_sub(void* r, const void* a, const void* b) { LD(fr, x, a); LD(fr, y, b); fr z; fr::sub(z, x, y); ST(r, z); }
void hs_fr_from_mont(void* r, const void* a) { LD(fr, x, a); uint32_t o[8]; fr_from_mont(o, x); memcpy(r, o, 32); }
void hs_fp2_mul(void* r, const void* a, const void* b) { LD(fp2, x, a); LD(fp2, y, b); fp2::mul(x, x, y); ST(r, x); }
void hs_fp2_sqr(void* r, const void* a) { LD(fp2, x, a); fp2::sqr(x, x); ST(r, x); }
void hs_fp2_inv(void* r, const void* a) { LD(fp2, x, a); fp2::inv(x, x); ST(r, x); }
void hs_fp6_mul(void* r, const void* a, const void* b) { LD(fp6, x, a); LD(fp6, y, b); fp6::mul(x, x, y); ST(r, x); }
void hs_fp6_inv(void* r, const void* a) { LD(fp6, x, a); fp6::inv(x, x); ST(r, x); }
void hs_fp12_mul(void* r, const void* a, const void* b) { LD(fp12, x, a); LD(fp12, y, b); fp12::mul(x, x, y); ST(r, x); }
void hs_fp12_sqr(void* r, const void* a) { LD(fp12, x, a); fp12::sqr(x, x); ST(r, x); }
void hs_fp12_inv(void* r, const void* a) { LD(fp12, x, a); fp12::inv(x, x); ST(r, x); }
void hs_fp12_frob1(void* r, const void* a) { LD(fp12, x, a); fp12::frobenius<1>(x, x); ST(r, x); }
void hs_fp12_frob2(void* r, const void* a) { LD(fp12, x, a); fp12::frobenius<2>(x, x); ST(r, x); }
void hs_fp12_cyclo_sqr(void* r, const void* a) { LD(fp12, x, a); fp12::cyclotomic_sqr(x, x); ST(r, x); }
void hs_fp12_mul_by_014(void* r, const void* a, const void* c0, const void* c1, const void* c4) {
  LD(fp12, x, a); LD(fp2, p, c0); LD(fp2, q, c1); LD(fp2, s, c4); fp12::mul_by_014(x, x, p, q, s); ST(r, x); }

// group ops: affine in / affine out (identity = all-zero)
void hs_g1_add(void* r, const void* a, const void* b) {
  LD(g1_aff, p, a); LD(g1_aff, q, b); g1_jac j; j.from_affine(p); g1_jac::add_mixed(j, j, q); g1_aff o; g1_jac::to_affine(o, j); ST(r, o); }
void hs_g1_add_full(void* r, const void* a, const void* b) {
  LD(g1_aff, p, a); LD(g1_aff, q, b); g1_jac j, k; j.from_affine(p); k.from_affine(q);
  g1_jac::dbl(j, j); g1_jac::dbl(k, k); g1_jac::add(j, j, k); g1_aff o; g1_jac::to_affine(o, j); ST(r, o); }  // 2a + 2b
void hs_g1_mul(void* r, const void* a, const void* k_mont) {
  LD(g1_aff, p, a); LD(fr, k, k_mont); uint32_t kk[8]; fr_from_mont(kk, k); g1_jac j; scalar_mul<FpOps>(j, p, kk); g1_aff o; g1_jac::to_affine(o, j); ST(r, o); }
void hs_g2_add(void* r, const void* a, const void* b) {
  LD(g2_aff, p, a); LD(g2_aff, q, b); g2_jac j; j.from_affine(p); g2_jac::add_mixed(j, j, q); g2_aff o; g2_jac::to_affine(o, j); ST(r, o); }
void hs_g2_add_full(void* r, const void* a, const void* b) {
  LD(g2_aff, p, a); LD(g2_aff, q, b); g2_jac j, k; j.from_affine(p); k.from_affine(q);
  g2_jac::dbl(j, j); g2_jac::dbl(k, k); g2_jac::add(j, j, k); g2_aff o; g2_jac::to_affine(o, j); ST(r, o); }
void hs_g2_mul(void* r, const void* a, const void* k_mont) {
  LD(g2_aff, p, a); LD(fr, k, k_mont); uint32_t kk[8]; fr_from_mont(kk, k); g2_jac j; scalar_mul<Fp2Ops>(j, p, kk); g2_aff o; g2_jac::to_affine(o, j); ST(r, o); }

// Miller product over n (G1,G2) affine pairs (identity pairs dropped), NO final exponentiation
void hs_miller(void* r, int n, const void* g1s, const void* g2s) {
  const g1_aff* P = (const g1_aff*)g1s; const g2_aff* Q = (const g2_aff*)g2s;
  std::vector<std::vector<line_coeffs>> lines; std::vector<g1_aff> ps;
  for (int i = 0; i < n; i++) {
    if (P[i].is_inf() || Q[i].is_inf()) continue;
    lines.emplace_back(GS_NUM_LINES); g2_prepare((uint32_t*)lines.back().data(), 1, Q[i]); ps.push_back(P[i]);
  }
  fp12 f; f.set_one(); int idx = 0;
  for (int b = 62; b >= 0; b--) {
    fp12::sqr(f, f);
    for (size_t k = 0; k < ps.size(); k++) miller_apply_line(f, lines[k][idx], ps[k].x, ps[k].y);
    idx++;
    if ((GS_X_ABS >> b) & 1) { for (size_t k = 0; k < ps.size(); k++) miller_apply_line(f, lines[k][idx], ps[k].x, ps[k].y); idx++; }
  }
  fp12::conj(f, f); ST(r, f);
}
void hs_final_exp(void* r, const void* a) { LD(fp12, x, a); fp12 o; final_exponentiation(o, x); ST(r, o); }
}

// ---- v2 Miller accumulator (w-basis, strided): same inputs/outputs as hs_miller
#include "../../tools/experimental/miller_v2.cuh"
extern "C" void hs_miller_v2(void* r, int n, const void* g1s, const void* g2s, int stride) {
  const g1_aff* P = (const g1_aff*)g1s; const g2_aff* Q = (const g2_aff*)g2s;
  std::vector<std::vector<line_coeffs>> lines; std::vector<g1_aff> ps;
  for (int i = 0; i < n; i++) {
    if (P[i].is_inf() || Q[i].is_inf()) continue;
    lines.emplace_back(GS_NUM_LINES); g2_prepare((uint32_t*)lines.back().data(), 1, Q[i]); ps.push_back(P[i]);
  }
  std::vector<uint32_t> f(144 * stride, 0xdeadbeef), lc(72 * stride, 0xdeadbeef);
  f12w_set_one(f.data(), stride);
  int idx = 0;
  for (int b = 62; b >= 0; b--) {
    f12w_sqr(f.data(), lc.data(), stride);
    int nl = ((GS_X_ABS >> b) & 1) ? 2 : 1;
    for (int t = 0; t < nl; t++, idx++)
      for (size_t k = 0; k < ps.size(); k++) {
        fp2 c1, c2;
        fp2::mul_fp(c1, lines[k][idx].c1, ps[k].x);
        fp2::mul_fp(c2, lines[k][idx].c2, ps[k].y);
        st_fp2(GS_COEF(lc.data(), 0, stride), stride, lines[k][idx].c0);
        st_fp2(GS_COEF(lc.data(), 1, stride), stride, c1);
        st_fp2(GS_COEF(lc.data(), 2, stride), stride, c2);
        f12w_mul_line(f.data(), lc.data(), stride);
      }
  }
  fp12 out; f12w_store_conj(out, f.data(), stride); ST(r, out);
}

// ---- v3 cooperative Miller accumulation (coop12.cuh): `lanes` independent accumulators, n pairs each
// (g1s/g2s indexed [lane*n + i]); the (k, lane) functions are run for all k, lane between "barriers".
#include "../../groth-sahai-rs_b200/csrc/coop12.cuh"
extern "C" void hs_miller_v3(void* r, int lanes, int n, const void* g1s, const void* g2s) {
  const g1_aff* P = (const g1_aff*)g1s; const g2_aff* Q = (const g2_aff*)g2s;
  const int TILE = 6 * CQ_FP;
  std::vector<uint32_t> tiles((size_t)n * GS_NUM_LINES * TILE, 0xdeadbeefu);
  std::vector<uint32_t> mask(n, 0);
  for (int lane = 0; lane < lanes; lane++)
    for (int i = 0; i < n; i++) {
      const g1_aff& p = P[lane * n + i]; const g2_aff& q = Q[lane * n + i];
      if (p.is_inf() || q.is_inf()) continue;
      mask[i] |= 1u << lane;
      g2_proj t; t.x = q.x; t.y = q.y; t.z.set_one();
      int idx = 0;
      for (int bit = 62; bit >= 0; bit--) {
        int nl = ((GS_X_ABS >> bit) & 1) ? 2 : 1;
        for (int w = 0; w < nl; w++, idx++) {
          line_coeffs l;
          if (w == 0) g2_double_step(t, l); else g2_add_step(t, q, l);
          uint32_t* o = tiles.data() + ((size_t)i * GS_NUM_LINES + idx) * TILE;
          fp v;
          cq_st(cq_ptr(o, 0, lane), l.c0.c0); cq_st(cq_ptr(o, 1, lane), l.c0.c1);
          fp::mul(v, l.c1.c0, p.x); cq_st(cq_ptr(o, 2, lane), v);
          fp::mul(v, l.c1.c1, p.x); cq_st(cq_ptr(o, 3, lane), v);
          fp::mul(v, l.c2.c0, p.y); cq_st(cq_ptr(o, 4, lane), v);
          fp::mul(v, l.c2.c1, p.y); cq_st(cq_ptr(o, 5, lane), v);
        }
      }
    }
  std::vector<uint32_t> acc(2 * CQ_ACC, 0xdeadbeefu);
  int cur = 0;
  for (int k = 0; k < 6; k++) for (int lane = 0; lane < 32; lane++) cq_set_one(k, lane, acc.data());
  int idx = 0;
  for (int bit = 62; bit >= 0; bit--) {
    if (bit != 62) {
      for (int k = 0; k < 6; k++) for (int lane = 0; lane < lanes; lane++) cq_sqr(k, lane, acc.data() + cur * CQ_ACC, acc.data() + (cur ^ 1) * CQ_ACC);
      cur ^= 1;
    }
    int nl = ((GS_X_ABS >> bit) & 1) ? 2 : 1;
    for (int w = 0; w < nl; w++, idx++)
      for (int i = 0; i < n; i++) {
        if (!mask[i]) continue;
        const uint32_t* tile = tiles.data() + ((size_t)i * GS_NUM_LINES + idx) * TILE;
        for (int k = 0; k < 6; k++) for (int lane = 0; lane < lanes; lane++)
          cq_line_mul(k, lane, acc.data() + cur * CQ_ACC, acc.data() + (cur ^ 1) * CQ_ACC, tile, (mask[i] >> lane) & 1);
        cur ^= 1;
      }
  }
  fp12* out = (fp12*)r;
  for (int lane = 0; lane < lanes; lane++)
    for (int k = 0; k < 6; k++) {
      fp2 c; cq_ld_coef(c.c0, c.c1, acc.data() + cur * CQ_ACC, k, lane, false, false);
      if (k & 1) fp2::neg(c, c);
      ((fp2*)&out[lane])[cq_tower_pos(k)] = c;
    }
}
// fout = f * g and f^2 through the cooperative ops (single lane), tower-ordered in/out
static void cq_from_tower(uint32_t* acc, int lane, const fp12& x) {
  for (int k = 0; k < 6; k++) cq_st_coef(acc, k, lane, ((const fp2*)&x)[cq_tower_pos(k)]);
}
static void cq_to_tower(fp12& x, const uint32_t* acc, int lane) {
  for (int k = 0; k < 6; k++) { fp2 c; cq_ld_coef(c.c0, c.c1, acc, k, lane, false, false); ((fp2*)&x)[cq_tower_pos(k)] = c; }
}
extern "C" void hs_cq_mul(void* r, const void* a, const void* b, int lane) {
  LD(fp12, x, a); LD(fp12, y, b);
  std::vector<uint32_t> f(CQ_ACC), g(CQ_ACC), o(CQ_ACC);
  cq_from_tower(f.data(), lane, x); cq_from_tower(g.data(), lane, y);
  for (int k = 0; k < 6; k++) cq_mul(k, lane, f.data(), g.data(), o.data());
  fp12 z; cq_to_tower(z, o.data(), lane); ST(r, z);
}
extern "C" void hs_cq_sqr(void* r, const void* a, int lane) {
  LD(fp12, x, a);
  std::vector<uint32_t> f(CQ_ACC), o(CQ_ACC);
  cq_from_tower(f.data(), lane, x);
  for (int k = 0; k < 6; k++) cq_sqr(k, lane, f.data(), o.data());
  fp12 z; cq_to_tower(z, o.data(), lane); ST(r, z);
}
// the cooperative final-exponentiation op program (finalexp.cu runs the same program on the GPU)
extern "C" void hs_final_exp3(void* r, const void* a, int lane) {
  LD(fp12, x, a);
  std::vector<uint32_t> bufs(CQ_FE_NBUF * CQ_ACC, 0xdeadbeefu);
  cq_from_tower(bufs.data(), lane, x);
  static uint32_t prog[CQ_FE_MAXOPS];
  int n = cq_build_final_exp(prog);
  for (int i = 0; i < n; i++)
    for (int k = 0; k < 6; k++) cq_exec(prog[i], k, lane, bufs.data());
  fp12 z; cq_to_tower(z, bufs.data() + CQ_FE_OUT * CQ_ACC, lane); ST(r, z);
}
extern "C" void hs_cq_cyc_sqr(void* r, const void* a, int lane) {
  LD(fp12, x, a);
  std::vector<uint32_t> f(CQ_ACC), o(CQ_ACC);
  cq_from_tower(f.data(), lane, x);
  for (int k = 0; k < 6; k++) cq_cyc_sqr(k, lane, f.data(), o.data());
  fp12 z; cq_to_tower(z, o.data(), lane); ST(r, z);
}
// safegcd inversion (modinv.cuh)
extern "C" void hs_fp_inv_sg(void* r, const void* a) { LD(fp, x, a); fp z; fp_inv_sg(z, x); ST(r, z); }

// ---- v4: affine line walk (4 points per inversion, safegcd) + unit-gamma line multiplication.
// One accumulator (lane), n <= 4 pairs.  The Miller value differs from hs_miller by Fp2 / Fp factors, so the
// caller compares after the final exponentiation.
extern "C" void hs_miller_v4(void* r, int n, const void* g1s, const void* g2s, int lane) {
  const g1_aff* P = (const g1_aff*)g1s; const g2_aff* Q = (const g2_aff*)g2s;
  const int TILE = 4 * CQ_FP;
  fp2 Tx[4], Ty[4], Qx[4], Qy[4], lam[4], mu[4]; bool act[4] = {false, false, false, false};
  fp s[4], w[4];
  for (int i = 0; i < n && i < 4; i++) {
    if (P[i].is_inf() || Q[i].is_inf()) continue;
    act[i] = true; Tx[i] = Q[i].x; Ty[i] = Q[i].y; Qx[i] = Q[i].x; Qy[i] = Q[i].y;
    fp_inv(w[i], P[i].y); fp::mul(s[i], P[i].x, w[i]); fp::neg(s[i], s[i]);
  }
  std::vector<uint32_t> acc(2 * CQ_ACC, 0xdeadbeefu), tile(TILE, 0xdeadbeefu);
  int cur = 0;
  for (int k = 0; k < 6; k++) cq_set_one(k, lane, acc.data());
  for (int bit = 62; bit >= 0; bit--) {
    if (bit != 62) {
      for (int k = 0; k < 6; k++) cq_sqr(k, lane, acc.data() + cur * CQ_ACC, acc.data() + (cur ^ 1) * CQ_ACC);
      cur ^= 1;
    }
    int nl = ((GS_X_ABS >> bit) & 1) ? 2 : 1;
    for (int t = 0; t < nl; t++) {
      g2_pts_arr TT{Tx, Ty}, QQ{Qx, Qy};
      g2_affine_step<4>(TT, QQ, act, t == 1, [&](int i, const fp2& l, const fp2& m) { lam[i] = l; mu[i] = m; }, [] {});
      for (int i = 3; i >= 0; i--) {
        if (!act[i]) continue;
        fp v;
        fp::mul(v, mu[i].c0, w[i]); cq_st(cq_ptr(tile.data(), 0, lane), v);
        fp::mul(v, mu[i].c1, w[i]); cq_st(cq_ptr(tile.data(), 1, lane), v);
        fp::mul(v, lam[i].c0, s[i]); cq_st(cq_ptr(tile.data(), 2, lane), v);
        fp::mul(v, lam[i].c1, s[i]); cq_st(cq_ptr(tile.data(), 3, lane), v);
        for (int k = 0; k < 6; k++) cq_line_mul_u(k, lane, acc.data() + cur * CQ_ACC, acc.data() + (cur ^ 1) * CQ_ACC, tile.data(), true);
        cur ^= 1;
      }
    }
  }
  fp12 out;
  for (int k = 0; k < 6; k++) {
    fp2 c; cq_ld_coef(c.c0, c.c1, acc.data() + cur * CQ_ACC, k, lane, false, false);
    if (k & 1) fp2::neg(c, c);
    ((fp2*)&out)[cq_tower_pos(k)] = c;
  }
  ST(r, out);
}

// FP64-pipe sum of products (fpd.cuh): the host stand-in for __fma_rz is fma() under FE_TOWARDZERO
extern "C" void hs_fp_mulsum_dfma(void* r, int nt, const void* a, const void* b) {
  const fp* A = (const fp*)a; const fp* B = (const fp*)b; fp z;
  const int old = fegetround();
  fesetround(FE_TOWARDZERO);
#define MS(NT) case NT: { fp x[NT], y[NT]; for (int i = 0; i < NT; i++) { x[i] = A[i]; y[i] = B[i]; } mulsum_dfma<NT>(z, x, y); \
    fp z2; mulsum_dfma_rolled<NT>(z2, x, [&](int t, int w) { return (uint64_t)y[t].l[w]; }); if (!z2.equals(z)) z.set_zero(); } break;
  switch (nt) { MS(1) MS(2) MS(3) MS(4) MS(6) MS(8) default: z.set_zero(); }
#undef MS
  fesetround(old);
  ST(r, z); }

// endomorphism splittings (endo.cuh): the decompositions and the per-part scalar multiplications
extern "C" void hs_glv_split(void* k1k2 /* 2 x 16 B */, const void* k_mont) {
  LD(fr, k, k_mont); uint32_t kk[8], a[4], b[4]; fr_from_mont(kk, k); glv_split(a, b, kk); memcpy(k1k2, a, 16); memcpy((char*)k1k2 + 16, b, 16); }
extern "C" void hs_gls_split(void* c /* 4 x 8 B */, const void* k_mont) {
  LD(fr, k, k_mont); uint32_t kk[8]; uint64_t d[4]; fr_from_mont(kk, k); gls_split(d, kk); memcpy(c, d, 32); }
extern "C" void hs_endo_psi(void* r, const void* a) { LD(g2_aff, p, a); g2_aff q; endo_psi(q, p); ST(r, q); }
// sum over the parts of EndoSplit<F>::part == k * base
extern "C" void hs_g1_mul_split(void* r, const void* a, const void* k_mont) {
  LD(g1_aff, p, a); LD(fr, k, k_mont); uint32_t kk[8]; fr_from_mont(kk, k); g1_jac acc; acc.set_inf();
  for (int j = 0; j < EndoSplit<FpOps>::PARTS; j++) { g1_jac t; EndoSplit<FpOps>::part(t, p, kk, j); g1_jac::add(acc, acc, t); }
  g1_aff o; g1_jac::to_affine(o, acc); ST(r, o); }
extern "C" void hs_g2_mul_split(void* r, const void* a, const void* k_mont) {
  LD(g2_aff, p, a); LD(fr, k, k_mont); uint32_t kk[8]; fr_from_mont(kk, k); g2_jac acc; acc.set_inf();
  for (int j = 0; j < EndoSplit<Fp2Ops>::PARTS; j++) { g2_jac t; EndoSplit<Fp2Ops>::part(t, p, kk, j); g2_jac::add(acc, acc, t); }
  g2_aff o; g2_jac::to_affine(o, acc); ST(r, o); }

// wire formats (wire.cuh): the one-point (de)compression the serialisation kernels run
extern "C" void hs_g1_compress(void* out48, const void* a) { LD(g1_aff, p, a); g1_compress_point((uint8_t*)out48, p); }
extern "C" int hs_g1_decompress(void* out, const void* in48, int check) { g1_aff p; bool ok = g1_decompress_point(p, (const uint8_t*)in48, check); ST(out, p); return ok; }
extern "C" void hs_g2_compress(void* out96, const void* a) { LD(g2_aff, p, a); g2_compress_point((uint8_t*)out96, p); }
extern "C" int hs_g2_decompress(void* out, const void* in96, int check) { g2_aff p; bool ok = g2_decompress_point(p, (const uint8_t*)in96, check); ST(out, p); return ok; }
extern "C" int hs_fp2_sqrt(void* r, const void* a) { LD(fp2, x, a); fp2 y; bool ok = fp2_sqrt(y, x); ST(r, y); return ok; }

// ---- randfold.cuh: the folding primitives of gs_verify_batch_rand
#include "randfold.cuh"
extern "C" void hs_rand_jsf(void* out /* 34 + 34 digits, then len as one byte */, uint32_t a, uint32_t b) {
  jsf33 j = make_jsf(a, b);
  memcpy(out, j.u0, 34);
  memcpy((char*)out + 34, j.u1, 34);
  ((uint8_t*)out)[68] = (uint8_t)j.len;
}
extern "C" void hs_rand_fold_g2(void* r, const void* y0, const void* y1, uint64_t w) {
  LD(g2_aff, a, y0); LD(g2_aff, b, y1);
  g2_aff o;
  rand_fold_g2_single(o, a, b, make_jsf((uint32_t)w, (uint32_t)(w >> 32)));
  ST(r, o);
}
extern "C" void hs_rand_fold_g1(void* r, const void* x0, const void* x1, uint64_t sg, uint64_t tu) {
  LD(g1_aff, a, x0); LD(g1_aff, b, x1);
  g1_aff o;
  rand_fold_g1_single(o, a, b, sg, tu);
  ST(r, o);
}
