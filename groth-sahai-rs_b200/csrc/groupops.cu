// Entry-wise arithmetic of the commitment groups and of Matrix<Fr>, batched:
//   Com1 / Com2  Add, Sub, Neg, Sum            src/data_structures.rs:162-255 (impl_base_commit_groups!)
//   ComT         Add (= GT product), Sub, Neg (= conjugate), Sum, Zero        :391-479
//   Matrix<Fr>   add, neg, scalar_mul (element-wise parts of the Mat trait)   :768-823
// One thread per affine point / GT value / scalar; point additions are normalised with one field inversion per
// block (batchinv.cuh) where the reference pays one per addition (:187-188).
#include "batchinv.cuh"
#include "ctx.h"
#include "prover_impl.cuh"  // reduce_rows, k_jac_rows_to_affine

using namespace gs;

namespace gs {

enum { OP_ADD = 0, OP_SUB = 1, OP_NEG = 2, OP_MUL = 3 };

// out[i] = a[i] (+|-) b[i]  or  -a[i]   over affine points (the identity is the all-zero encoding)
template <class F>
__global__ void __launch_bounds__(128) k_aff_binop(const Aff<F>* __restrict__ a, const Aff<F>* __restrict__ b, Aff<F>* __restrict__ out,
                                                   size_t n, int op) {
  __shared__ fp sm[2 * 128];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = i < n;
  Jac<F> acc;
  acc.set_inf();
  if (active) {
    Aff<F> p = a[i];
    if (op == OP_NEG) {
      if (!p.is_inf()) F::neg(p.y, p.y);
      acc.from_affine(p);
    } else {
      Aff<F> q = b[i];
      if (op == OP_SUB && !q.is_inf()) F::neg(q.y, q.y);
      acc.from_affine(p);
      Jac<F>::add_mixed(acc, acc, q);
    }
  }
  Aff<F> r;
  block_to_affine<128>(r, acc, sm);
  if (active) out[i] = r;
}

// terms[c][t] = Jacobian form of a[t*C + c]: column c of an array of n elements with C points each (Com = 2 points)
template <class F>
__global__ void k_aff_to_jac_cols(const Aff<F>* __restrict__ a, Jac<F>* __restrict__ terms, size_t n, int C) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n * C) return;
  size_t t = id % n;
  int c = (int)(id / n);
  Jac<F> j;
  j.from_affine(a[t * C + c]);
  terms[(size_t)c * n + t] = j;
}

// out[i] = a[i] * b[i] | a[i] * conj(b[i]) | conj(a[i])   in GT written additively (PairingOutput)
__global__ void __launch_bounds__(64) k_gt_binop(const fp12* __restrict__ a, const fp12* __restrict__ b, fp12* __restrict__ out, size_t n,
                                                 int op) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fp12 x = a[i], r;
  if (op == OP_NEG) {
    fp12::conj(r, x);
  } else {
    fp12 y = b[i];
    if (op == OP_SUB) fp12::conj(y, y);
    fp12::mul(r, x, y);
  }
  out[i] = r;
}
// in-place pairwise tree step over ComT entries: g[t] *= g[t + half]   (t + half < cur), g = [n][4] GT values
__global__ void __launch_bounds__(64) k_gt_reduce_step(fp12* __restrict__ g, size_t cur, size_t half) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= half * 4) return;
  size_t t = id >> 2;
  int e = (int)(id & 3);
  if (t + half >= cur) return;
  fp12 x = g[t * 4 + e], y = g[(t + half) * 4 + e], r;
  fp12::mul(r, x, y);
  g[t * 4 + e] = r;
}

// Matrix<Fr> element-wise: out = a + b | a - b | -a | s * a (s = b[0])
__global__ void k_fr_binop(const fr* __restrict__ a, const fr* __restrict__ b, fr* __restrict__ out, size_t n, int op) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fr x = a[i], r;
  if (op == OP_ADD)
    fr::add(r, x, b[i]);
  else if (op == OP_SUB)
    fr::sub(r, x, b[i]);
  else if (op == OP_NEG)
    fr::neg(r, x);
  else
    fr::mul(r, x, b[0]);
  out[i] = r;
}

}  // namespace gs

namespace {

template <class F>
int aff_binop(gs_ctx* ctx, size_t npts, const void* a, const void* b, void* out, int op) {
  if (!ctx || !a || !out || (op != OP_NEG && !b)) return GS_EARG;
  if (npts == 0) return GS_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  Aff<F>*da, *db = nullptr, *dout;
  CUDA_TRY(upload(ctx, sc, &da, a, npts));
  if (op != OP_NEG) CUDA_TRY(upload(ctx, sc, &db, b, npts));
  CUDA_TRY(sc.alloc(&dout, npts));
  LAUNCH((k_aff_binop<F>), npts, da, db, dout, npts, op);
  CUDA_TRY(cudaMemcpyAsync(out, dout, npts * sizeof(Aff<F>), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

// Sum for Com1 / Com2 (data_structures.rs:245-250): fold from zero => an empty input gives the identity pair
template <class F>
int com_sum(gs_ctx* ctx, size_t n, const void* a, void* out) {
  if (!ctx || !out || (n && !a)) return GS_EARG;
  if (n == 0) {
    memset(out, 0, 2 * sizeof(Aff<F>));
    return GS_OK;
  }
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  Aff<F>*da, *dout;
  Jac<F>* terms;
  CUDA_TRY(upload(ctx, sc, &da, a, 2 * n));
  CUDA_TRY(sc.alloc(&terms, 2 * n));
  CUDA_TRY(sc.alloc(&dout, 2));
  LAUNCH((k_aff_to_jac_cols<F>), 2 * n, da, terms, n, 2);
  int rc = gsi::reduce_rows<F>(ctx, terms, n, n, 2);
  if (rc) return rc;
  LAUNCH((k_jac_rows_to_affine<F>), (size_t)2, dout, terms, n, (size_t)2);
  CUDA_TRY(cudaMemcpyAsync(out, dout, 2 * sizeof(Aff<F>), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

int gt_binop(gs_ctx* ctx, size_t n, const void* a, const void* b, void* out, int op) {
  if (!ctx || !a || !out || (op != OP_NEG && !b)) return GS_EARG;
  if (n == 0) return GS_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  fp12 *da, *db = nullptr, *dout;
  CUDA_TRY(upload(ctx, sc, &da, a, n));
  if (op != OP_NEG) CUDA_TRY(upload(ctx, sc, &db, b, n));
  CUDA_TRY(sc.alloc(&dout, n));
  LAUNCH_CFG(k_gt_binop, n, 64, 0, da, db, dout, n, op);
  CUDA_TRY(cudaMemcpyAsync(out, dout, n * sizeof(fp12), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

int fr_binop(gs_ctx* ctx, size_t n, const gs_fr* a, const gs_fr* b, size_t nb, gs_fr* out, int op) {
  if (!ctx || !a || !out || (op != OP_NEG && !b)) return GS_EARG;
  if (n == 0) return GS_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  fr *da, *db = nullptr, *dout;
  CUDA_TRY(upload(ctx, sc, &da, a, n));
  if (op != OP_NEG) CUDA_TRY(upload(ctx, sc, &db, b, nb));
  CUDA_TRY(sc.alloc(&dout, n));
  LAUNCH(k_fr_binop, n, da, db, dout, n, op);
  CUDA_TRY(cudaMemcpyAsync(out, dout, n * sizeof(fr), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

}  // namespace

extern "C" {

int gs_com1_add(gs_ctx* ctx, size_t n, const gs_com1* a, const gs_com1* b, gs_com1* out) { return aff_binop<FpOps>(ctx, 2 * n, a, b, out, OP_ADD); }
int gs_com1_sub(gs_ctx* ctx, size_t n, const gs_com1* a, const gs_com1* b, gs_com1* out) { return aff_binop<FpOps>(ctx, 2 * n, a, b, out, OP_SUB); }
int gs_com1_neg(gs_ctx* ctx, size_t n, const gs_com1* a, gs_com1* out) { return aff_binop<FpOps>(ctx, 2 * n, a, nullptr, out, OP_NEG); }
int gs_com1_sum(gs_ctx* ctx, size_t n, const gs_com1* a, gs_com1* out) { return com_sum<FpOps>(ctx, n, a, out); }
int gs_com2_add(gs_ctx* ctx, size_t n, const gs_com2* a, const gs_com2* b, gs_com2* out) { return aff_binop<Fp2Ops>(ctx, 2 * n, a, b, out, OP_ADD); }
int gs_com2_sub(gs_ctx* ctx, size_t n, const gs_com2* a, const gs_com2* b, gs_com2* out) { return aff_binop<Fp2Ops>(ctx, 2 * n, a, b, out, OP_SUB); }
int gs_com2_neg(gs_ctx* ctx, size_t n, const gs_com2* a, gs_com2* out) { return aff_binop<Fp2Ops>(ctx, 2 * n, a, nullptr, out, OP_NEG); }
int gs_com2_sum(gs_ctx* ctx, size_t n, const gs_com2* a, gs_com2* out) { return com_sum<Fp2Ops>(ctx, n, a, out); }

int gs_comt_add(gs_ctx* ctx, size_t n, const gs_comt* a, const gs_comt* b, gs_comt* out) { return gt_binop(ctx, 4 * n, a, b, out, OP_ADD); }
int gs_comt_sub(gs_ctx* ctx, size_t n, const gs_comt* a, const gs_comt* b, gs_comt* out) { return gt_binop(ctx, 4 * n, a, b, out, OP_SUB); }
int gs_comt_neg(gs_ctx* ctx, size_t n, const gs_comt* a, gs_comt* out) { return gt_binop(ctx, 4 * n, a, nullptr, out, OP_NEG); }
int gs_comt_sum(gs_ctx* ctx, size_t n, const gs_comt* a, gs_comt* out) {
  if (!ctx || !out || (n && !a)) return GS_EARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  fp12* d;
  if (n == 0) {  // ComT::zero(): four GT identities
    fp12 one[4];
    for (int e = 0; e < 4; e++) one[e].set_one();
    memcpy(out, one, sizeof(one));
    return GS_OK;
  }
  CUDA_TRY(upload(ctx, sc, &d, a, 4 * n));
  size_t cur = n;
  while (cur > 1) {
    size_t half = (cur + 1) / 2;
    LAUNCH_CFG(k_gt_reduce_step, half * 4, 64, 0, d, cur, half);
    cur = half;
  }
  CUDA_TRY(cudaMemcpyAsync(out, d, 4 * sizeof(fp12), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

int gs_fr_add(gs_ctx* ctx, size_t n, const gs_fr* a, const gs_fr* b, gs_fr* out) { return fr_binop(ctx, n, a, b, n, out, OP_ADD); }
int gs_fr_sub(gs_ctx* ctx, size_t n, const gs_fr* a, const gs_fr* b, gs_fr* out) { return fr_binop(ctx, n, a, b, n, out, OP_SUB); }
int gs_fr_neg(gs_ctx* ctx, size_t n, const gs_fr* a, gs_fr* out) { return fr_binop(ctx, n, a, nullptr, 0, out, OP_NEG); }
int gs_fr_scale(gs_ctx* ctx, size_t n, const gs_fr* s, const gs_fr* a, gs_fr* out) { return fr_binop(ctx, n, a, s, 1, out, OP_MUL); }

}  // extern "C"
