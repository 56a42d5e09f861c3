// Prover-side kernels: CRS generation, fixed-base window tables + batch commitments,
// Fr matrix algebra, variable-base MSMs for the proof elements, Mat products on Com1/Com2.
//
// Replaces (reference): src/generator.rs:81-118, src/prover/commit.rs:78-256,
// src/prover/prove.rs:92-488, src/data_structures.rs:645-742 / 768-913.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace gs {

// ------------------------------------------------------------------ small helpers
__global__ void k_fp12_set_one(fp12* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i].set_one();
}
// iota_T for PPE: (1, 1, 1, t)            data_structures.rs:509-516
__global__ void k_linear_map_ppe(const fp12* t, fp12* out) {
  int e = threadIdx.x;
  if (e >= 4) return;
  if (e == 3)
    out[3] = *t;
  else
    out[e].set_one();
}
// iota_T for the other three types as ONE (Com1, Com2) pair          data_structures.rs:519-540
//   MSMEG1: F(iota_1(t), W2)   MSMEG2: F(W1, iota_2(t))   Quad: F(W1, t W2) = F(t W1, W2)
__global__ void k_linear_map_slots(int type, const void* target, const crs_dev* crs, g1_aff* X, g2_aff* Y) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (type == 1) {
    X[0].set_inf();
    X[1] = *(const g1_aff*)target;
    Y[0] = crs->w2[0];
    Y[1] = crs->w2[1];
  } else if (type == 2) {
    X[0] = crs->w1[0];
    X[1] = crs->w1[1];
    Y[0].set_inf();
    Y[1] = *(const g2_aff*)target;
  } else {
    uint32_t k[8];
    fr_from_mont(k, *(const fr*)target);
    for (int a = 0; a < 2; a++) {
      g1_jac j;
      scalar_mul<FpOps>(j, crs->w1[a], k);
      g1_jac::to_affine(X[a], j);
    }
    Y[0] = crs->w2[0];
    Y[1] = crs->w2[1];
  }
}

// ------------------------------------------------------------------ CRS
struct crs_gen_in {
  g1_aff p1;
  g2_aff p2;
  fr a1, a2, t1, t2;
};
struct crs_gen_out {
  g1_aff p1, q1, u1, v1;
  g2_aff p2, q2, u2, v2;
};
// generator.rs:96-109: q1 = a1 p1, u1 = t1 p1, v1 = t1 q1 = (t1 a1) p1 ; same on G2.   6 threads.
__global__ void k_crs_generate(const crs_gen_in* in, crs_gen_out* out) {
  int t = threadIdx.x;
  if (blockIdx.x != 0 || t >= 6) return;
  fr s;
  if (t == 0 || t == 3) s = (t == 0) ? in->a1 : in->a2;
  if (t == 1 || t == 4) s = (t == 1) ? in->t1 : in->t2;
  if (t == 2) fr::mul(s, in->a1, in->t1);
  if (t == 5) fr::mul(s, in->a2, in->t2);
  uint32_t k[8];
  fr_from_mont(k, s);
  if (t < 3) {
    g1_jac j;
    scalar_mul<FpOps>(j, in->p1, k);
    g1_aff a;
    g1_jac::to_affine(a, j);
    if (t == 0) out->q1 = a;
    if (t == 1) out->u1 = a;
    if (t == 2) out->v1 = a;
    if (t == 0) out->p1 = in->p1;
  } else {
    g2_jac j;
    scalar_mul<Fp2Ops>(j, in->p2, k);
    g2_aff a;
    g2_jac::to_affine(a, j);
    if (t == 3) out->q2 = a;
    if (t == 4) out->u2 = a;
    if (t == 5) out->v2 = a;
    if (t == 3) out->p2 = in->p2;
  }
}

// W1 = u2 + (O, g1), W2 = v2 + (O, g2), and the negations used by verify
__global__ void k_crs_derive(crs_dev* c) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  c->w1[0] = c->u[1][0];
  {
    g1_jac j;
    j.from_affine(c->u[1][1]);
    g1_jac::add_mixed(j, j, c->g1);
    g1_jac::to_affine(c->w1[1], j);
  }
  c->w2[0] = c->v[1][0];
  {
    g2_jac j;
    j.from_affine(c->v[1][1]);
    g2_jac::add_mixed(j, j, c->g2);
    g2_jac::to_affine(c->w2[1], j);
  }
  for (int k = 0; k < 2; k++)
    for (int a = 0; a < 2; a++) {
      c->neg_u[k][a] = c->u[k][a];
      fp::neg(c->neg_u[k][a].y, c->neg_u[k][a].y);
    }
  for (int a = 0; a < 2; a++) {
    c->neg_w1[a] = c->w1[a];
    fp::neg(c->neg_w1[a].y, c->neg_w1[a].y);
  }
}

// ------------------------------------------------------------------ fixed-base window tables
// For base point B (one coordinate of u1, u2, W1 / v1, v2, W2) and window w:
//     T[w][d-1] = d * 2^(c w) * B,   d = 1 .. 2^(c-1)     (signed digits => half tables)
// layout: tab[((base*2 + a) * W + w) * H + (d-1)]
constexpr int GS_TAB_NT = 128;

template <class F>
__global__ void k_table_window_bases(const Aff<F>* __restrict__ bases, Aff<F>* __restrict__ tab, int c, int W, size_t H) {
  // one thread per base point: writes d = 1 entries (2^(cw) B) for every window
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= 6) return;
  Jac<F> j;
  j.from_affine(bases[b]);
  for (int w = 0; w < W; w++) {
    Aff<F> a;
    Jac<F>::to_affine(a, j);
    tab[((size_t)b * W + w) * H] = a;
    for (int i = 0; i < c; i++) Jac<F>::dbl(j, j);
  }
}

// thread -> (base b, window w, run r): entries d = r*RUN+1 .. r*RUN+RUN of T[b][w] by a chain of
// mixed additions from a small scalar-mul start, normalised with a block-wide batch inversion per step.
template <class F, int RUN>
__global__ void __launch_bounds__(GS_TAB_NT) k_table_fill(Aff<F>* __restrict__ tab, int W, size_t H) {
  __shared__ fp sm[2 * GS_TAB_NT];
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t runs_per_row = (H + RUN - 1) / RUN;
  size_t total = (size_t)6 * W * runs_per_row;
  bool active = id < total;
  size_t row = active ? id / runs_per_row : 0, r = active ? id % runs_per_row : 0;
  Aff<F>* T = tab + row * H;
  Aff<F> B = T[0];
  Jac<F> acc;
  uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  k[0] = (uint32_t)(r * RUN);  // start = (r*RUN) * B
  scalar_mul<F>(acc, B, k);
  for (int i = 0; i < RUN; i++) {
    Jac<F>::add_mixed(acc, acc, B);
    Aff<F> a;
    block_to_affine<GS_TAB_NT>(a, acc, sm);
    size_t d1 = r * RUN + i;  // d - 1
    if (active && d1 < H && d1 > 0) T[d1] = a;
  }
}

struct fixed_tables {
  int c = 0, W = 0;
  size_t H = 0;
  g1_aff* t1 = nullptr;  // bases: u1.0 u1.1 u2.0 u2.1 W1.0 W1.1
  g2_aff* t2 = nullptr;  // bases: v1.0 v1.1 v2.0 v2.1 W2.0 W2.1
  const crs_dev* crs = nullptr;
  cudaStream_t stream = nullptr;
  uint64_t* launches = nullptr;

  void release() {
    if (t1) cudaFree(t1);
    if (t2) cudaFree(t2);
    t1 = nullptr;
    t2 = nullptr;
    c = 0;
  }
  // called at CRS load: small (c = 8) tables, built in well under a millisecond of GPU time
  int build(cudaStream_t s, const crs_dev* crs_, uint64_t* launch_counter) {
    stream = s;
    crs = crs_;
    launches = launch_counter;
    release();
    return rebuild(8);
  }
  // larger windows are built lazily by the first big batch (see gs_batch_commit_*)
  int ensure(int want_c) { return want_c == c ? 0 : rebuild(want_c); }

  int rebuild(int new_c) {
    release();
    c = new_c;
    W = (256 + c - 1) / c;
    H = (size_t)1 << (c - 1);
    size_t n = (size_t)6 * W * H;
    if (cudaMalloc(&t1, n * sizeof(g1_aff)) != cudaSuccess) return 1;
    if (cudaMalloc(&t2, n * sizeof(g2_aff)) != cudaSuccess) return 1;
    // the six base points are contiguous in crs_dev in exactly the table order: u[2][2] then w1[2]
    g1_aff* b1;
    g2_aff* b2;
    if (cudaMallocAsync(&b1, 6 * sizeof(g1_aff), stream) != cudaSuccess) return 1;
    if (cudaMallocAsync(&b2, 6 * sizeof(g2_aff), stream) != cudaSuccess) return 1;
    cudaMemcpyAsync(b1, &crs->u[0][0], 4 * sizeof(g1_aff), cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(b1 + 4, &crs->w1[0], 2 * sizeof(g1_aff), cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(b2, &crs->v[0][0], 4 * sizeof(g2_aff), cudaMemcpyDeviceToDevice, stream);
    cudaMemcpyAsync(b2 + 4, &crs->w2[0], 2 * sizeof(g2_aff), cudaMemcpyDeviceToDevice, stream);
    k_table_window_bases<FpOps><<<1, 32, 0, stream>>>(b1, t1, c, W, H);
    k_table_window_bases<Fp2Ops><<<1, 32, 0, stream>>>(b2, t2, c, W, H);
    constexpr int RUN = 16;
    size_t threads = (size_t)6 * W * ((H + RUN - 1) / RUN);
    unsigned grid = (unsigned)((threads + GS_TAB_NT - 1) / GS_TAB_NT);
    k_table_fill<FpOps, RUN><<<grid, GS_TAB_NT, 0, stream>>>(t1, W, H);
    k_table_fill<Fp2Ops, RUN><<<grid, GS_TAB_NT, 0, stream>>>(t2, W, H);
    *launches += 4;
    cudaFreeAsync(b1, stream);
    cudaFreeAsync(b2, stream);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
  }
};

// ------------------------------------------------------------------ batch commitments
// out[i].p[a] = s0_i * Base0.a + s1_i * Base1.a  (+ addend_i when a == 1)
//   batch_commit_G1 (commit.rs:78-100):            s0,s1 = R[i][0], R[i][1]; bases u1,u2; addend X_i
//   batch_commit_scalar_to_B1 (commit.rs:125-156): s0 = x_i, s1 = r_i;     bases W1,u1; no addend
// thread -> (i, a); signed c-bit windows; one table lookup + one mixed addition per window.
// c bits of the 256-bit integer k starting at `bit` (zero beyond bit 255); c <= 16
GS_HD GS_INL uint32_t get_bits(const uint32_t k[8], int bit, int c) {
  int w = bit >> 5, s = bit & 31;
  if (w >= 8) return 0;
  uint64_t v = k[w];
  if (w + 1 < 8) v |= (uint64_t)k[w + 1] << 32;
  return (uint32_t)(v >> s) & ((1u << c) - 1u);
}

template <class F>
GS_HD GS_INL void fixed_base_accumulate(Jac<F>& acc, const Aff<F>* __restrict__ T /* [W][H] */, const uint32_t k[8], int c,
                                        int W, size_t H) {
  uint32_t carry = 0;
  const uint32_t half = 1u << (c - 1);
  for (int w = 0; w < W; w++) {
    uint32_t d = get_bits(k, w * c, c) + carry;   // digits recoded into (-2^(c-1), 2^(c-1)]
    bool negd = d > half;
    carry = negd ? 1u : 0u;
    uint32_t mag = negd ? (1u << c) - d : d;
    if (mag == 0) continue;
    Aff<F> e = T[(size_t)w * H + (mag - 1)];
    if (negd) F::neg(e.y, e.y);
    Jac<F>::add_mixed(acc, acc, e);
  }
}

template <class F>
__global__ void __launch_bounds__(GS_TAB_NT) k_fixed_commit(const Aff<F>* __restrict__ tab, int c, int W, size_t H, int base0,
                                                            int base1, const fr* __restrict__ s0, size_t s0_stride,
                                                            const fr* __restrict__ s1, size_t s1_stride,
                                                            const Aff<F>* __restrict__ addend, Aff<F>* __restrict__ out, size_t n) {
  __shared__ fp sm[2 * GS_TAB_NT];
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool active = id < 2 * n;
  size_t i = active ? id >> 1 : 0;
  int a = (int)(id & 1);
  Jac<F> acc;
  acc.set_inf();
  if (active) {
    uint32_t k[8];
    fr_from_mont(k, s0[i * s0_stride]);
    fixed_base_accumulate<F>(acc, tab + ((size_t)(base0 * 2 + a) * W) * H, k, c, W, H);
    fr_from_mont(k, s1[i * s1_stride]);
    fixed_base_accumulate<F>(acc, tab + ((size_t)(base1 * 2 + a) * W) * H, k, c, W, H);
    if (addend != nullptr && a == 1) Jac<F>::add_mixed(acc, acc, addend[i]);
  }
  Aff<F> r;
  block_to_affine<GS_TAB_NT>(r, acc, sm);
  if (active) out[i * 2 + a] = r;
}

// ------------------------------------------------------------------ Fr matrix algebra
// out (r x c) = A (r x k) * B (k x c), row-major; optional transposes via strides
__global__ void k_fr_matmul(fr* __restrict__ out, const fr* __restrict__ A, size_t a_rs, size_t a_cs, const fr* __restrict__ B,
                            size_t b_rs, size_t b_cs, size_t r, size_t k, size_t c) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= r * c) return;
  size_t i = id / c, j = id % c;
  fr acc;
  acc.set_zero();
  for (size_t t = 0; t < k; t++) {
    fr p;
    fr::mul(p, A[i * a_rs + t * a_cs], B[t * b_rs + j * b_cs]);
    fr::add(acc, acc, p);
  }
  out[id] = acc;
}

// coef_pi[i][l] = (RG * S)[i][l] - T[l][i]            (prove.rs:139-142)   cx x cy
__global__ void k_coef_pi(fr* __restrict__ out, const fr* __restrict__ RG, const fr* __restrict__ S, const fr* __restrict__ T, int cx,
                          int cy, size_t n) {
  int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= cx * cy) return;
  int i = id / cy, l = id % cy;
  fr acc;
  acc.set_zero();
  for (size_t j = 0; j < n; j++) {
    fr p;
    fr::mul(p, RG[i * n + j], S[j * cy + l]);
    fr::add(acc, acc, p);
  }
  fr::sub(acc, acc, T[l * cx + i]);
  out[id] = acc;
}

// scalar vectors of the variable-base part of a proof element:
//   sv[i][t] = R[t][i] (t < m) ; RG[i][t-m] (t >= m)          i < cx
__global__ void k_concat_scalars(fr* __restrict__ sv, const fr* __restrict__ R, const fr* __restrict__ RG, int cx, size_t m, size_t n) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (size_t)cx * (m + n)) return;
  size_t i = id / (m + n), t = id % (m + n);
  sv[id] = t < m ? R[t * cx + i] : RG[i * n + (t - m)];
}

// dot[i] = sum_t sv[i][t] * w[t]    (scalar-typed constants/variables: everything collapses onto W)
__global__ void k_fr_dot(fr* __restrict__ out, const fr* __restrict__ sv, const fr* __restrict__ w0, size_t m, const fr* __restrict__ w1,
                         size_t n, int rows) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  fr acc;
  acc.set_zero();
  for (size_t t = 0; t < m + n; t++) {
    fr p;
    fr::mul(p, sv[(size_t)i * (m + n) + t], t < m ? w0[t] : w1[t - m]);
    fr::add(acc, acc, p);
  }
  out[i] = acc;
}

// ------------------------------------------------------------------ variable-base MSM (proof elements)
// terms[row][t] = sv[row][t] * base[t]   (4-bit signed windows per term), bases = two concatenated segments
template <class F>
__global__ void __launch_bounds__(128) k_msm_terms(Jac<F>* __restrict__ terms, const fr* __restrict__ sv, const Aff<F>* __restrict__ b0,
                                                   size_t n0, const Aff<F>* __restrict__ b1, size_t n1, int rows) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t nt = n0 + n1;
  if (id >= nt * rows) return;
  size_t t = id % nt;
  Aff<F> B = t < n0 ? b0[t] : b1[t - n0];
  uint32_t k[8];
  fr_from_mont(k, sv[id]);
  Jac<F> j;
  scalar_mul<F>(j, B, k);
  terms[id] = j;
}

// in-place pairwise tree reduction: terms[row][t] += terms[row][t + half] for t < half (one launch per level)
template <class F>
__global__ void __launch_bounds__(128) k_jac_reduce_step(Jac<F>* __restrict__ terms, size_t row_stride, size_t cur, size_t half, int rows) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= half * rows) return;
  size_t row = id / half, t = id % half;
  if (t + half >= cur) return;
  Jac<F>* base = terms + row * row_stride;
  Jac<F> a = base[t], b = base[t + half];
  Jac<F>::add(a, a, b);
  base[t] = a;
}

// final assembly of a proof element (prove.rs:146, 162):
//   out[i].p[0] =                  sum_l coef[i][l] key_l.0  (+ e_i W.0)
//   out[i].p[1] = varsum[i]      + sum_l coef[i][l] key_l.1  (+ e_i W.1)
// thread -> (i, a).  `varsum` = reduced MSM rows (group-typed) or null; `e` = collapsed scalar (scalar-typed) or null.
template <class F>
__global__ void k_proof_finish(Aff<F>* __restrict__ out, int rows, int ncoef, const fr* __restrict__ coef, size_t coef_rs,
                               size_t coef_cs, const Aff<F>* __restrict__ key /* [2][2] */, const Jac<F>* __restrict__ varsum,
                               size_t var_stride, const fr* __restrict__ e, const Aff<F>* __restrict__ W /* [2] */) {
  int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= rows * 2) return;
  int i = id >> 1, a = id & 1;
  Jac<F> acc;
  acc.set_inf();
  if (varsum != nullptr && a == 1) acc = varsum[(size_t)i * var_stride];
  uint32_t k[8];
  for (int l = 0; l < ncoef; l++) {
    fr_from_mont(k, coef[i * coef_rs + l * coef_cs]);
    Jac<F> t;
    scalar_mul<F>(t, key[l * 2 + a], k);
    Jac<F>::add(acc, acc, t);
  }
  if (e != nullptr) {
    fr_from_mont(k, e[i]);
    Jac<F> t;
    scalar_mul<F>(t, W[a], k);
    Jac<F>::add(acc, acc, t);
  }
  Aff<F> r;
  Jac<F>::to_affine(r, acc);
  out[i * 2 + a] = r;
}

// ------------------------------------------------------------------ Mat::left_mul on Com matrices
// terms[(i*c + j)*2 + a][t] = lhs[i][t] * mat[t][j].a      (data_structures.rs:696-742)
template <class F>
__global__ void __launch_bounds__(128) k_com_matmul_terms(Jac<F>* __restrict__ terms, const fr* __restrict__ lhs,
                                                          const Aff<F>* __restrict__ mat, size_t r, size_t k, size_t c) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= r * c * 2 * k) return;
  size_t t = id % k;
  size_t o = id / k;
  int a = (int)(o & 1);
  size_t ij = o >> 1;
  size_t i = ij / c, j = ij % c;
  uint32_t kk[8];
  fr_from_mont(kk, lhs[i * k + t]);
  Jac<F> acc;
  scalar_mul<F>(acc, mat[(t * c + j) * 2 + a], kk);
  terms[id] = acc;
}
template <class F>
__global__ void k_jac_rows_to_affine(Aff<F>* __restrict__ out, const Jac<F>* __restrict__ terms, size_t row_stride, size_t rows) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= rows) return;
  Aff<F> r;
  Jac<F>::to_affine(r, terms[id * row_stride]);
  out[id] = r;
}

}  // namespace gs
