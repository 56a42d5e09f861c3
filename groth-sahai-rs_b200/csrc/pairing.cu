// Pairing-product pipeline: G2 line preparation, Miller accumulation, final exponentiation, and the
// ComT entry points of the C ABI (ComT::pairing / pairing_sum / linear_map_*, E::pairing).
// Reference: src/data_structures.rs:484-540, src/generator.rs:116.
#include "ctx.h"
#include "miller_v2.cuh"

using namespace gs;

namespace gs {

// "Slot" layout used by every pairing-product evaluation (ComT::pairing, pairing_sum, verify): a problem
// is K pairs (X_k in Com1, Y_k in Com2), and the wanted ComT is
//     ComT[a][b] = prod_k e(X_k.a, Y_k.b)          (src/data_structures.rs:494-502)
// Points are stored SoA over problems so that a warp (32 consecutive problems, same slot, same
// coordinate) reads contiguous memory:
//     X[(a*K + k) * nprob + p]   g1_aff        Y[(b*K + k) * nprob + p]   g2_aff
//     L[(((b*K + k) * 68 + step) * 72 + w) * nprob + p]   32-bit word w of a line triple (k_g2_prepare)

// ------------------------------------------------------------------ G2 preparation
// one thread per G2 point q = (b*K + k) * nprob + p
__global__ void __launch_bounds__(128) k_g2_prepare(const g2_aff* __restrict__ Y, uint32_t* __restrict__ L,
                                                    uint8_t* __restrict__ yinf, size_t npoints, size_t nprob) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npoints) return;
  g2_aff pt = Y[q];
  bool inf = pt.is_inf();
  yinf[q] = inf ? 1 : 0;
  if (inf) return;
  size_t bk = q / nprob, p = q % nprob;
  g2_prepare(L + (bk * GS_NUM_LINES * GS_LINE_WORDS) * nprob + p, nprob, pt);
}

// ------------------------------------------------------------------ Miller accumulation (v2)
// thread -> (p, e = 2a+b, chunk);  F[(chunk*4 + e) * nprob + p] = conj( prod over its slots ).
// The accumulator and the current line triple live in shared memory (miller_v2.cuh): 864 B / thread,
// 2 blocks of 128 threads per SM, no local-memory temporaries.
constexpr int GS_MV2_NT = 128;
constexpr int GS_MV2_SMEM = (144 + 72) * GS_MV2_NT * 4;
__global__ void __launch_bounds__(GS_MV2_NT, 2) k_miller(const g1_aff* __restrict__ X, const uint8_t* __restrict__ yinf,
                                                        const uint32_t* __restrict__ L, fp12* __restrict__ F, size_t nprob,
                                                        int K, int S, int nchunk) {
  extern __shared__ uint32_t sm[];
  uint32_t* f = sm + threadIdx.x;
  uint32_t* lc = sm + 144 * GS_MV2_NT + threadIdx.x;
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nprob * 4 * (size_t)nchunk) return;
  size_t p = id % nprob;
  int e = (int)((id / nprob) & 3);
  int ch = (int)(id / (nprob * 4));
  int a = e >> 1, b = e & 1;
  int k0 = ch * S, k1 = min(K, k0 + S);
  bool any = false;
  for (int k = k0; k < k1; k++) {
    if (yinf[((size_t)b * K + k) * nprob + p]) continue;
    if (X[((size_t)a * K + k) * nprob + p].is_inf()) continue;
    any = true;
  }
  fp12 out;
  if (!any) {
    out.set_one();
  } else {
    f12w_set_one(f, GS_MV2_NT);
    int idx = 0;
    for (int bit = 62; bit >= 0; bit--) {
      if (bit != 62) f12w_sqr(f, lc, GS_MV2_NT);  // f = 1 on the first pass
      int nl = ((GS_X_ABS >> bit) & 1) ? 2 : 1;
      for (int t = 0; t < nl; t++, idx++) {
        for (int k = k0; k < k1; k++) {
          size_t bk = (size_t)b * K + k;
          if (yinf[bk * nprob + p]) continue;
          const g1_aff* P = &X[((size_t)a * K + k) * nprob + p];
          fp px = P->x, py = P->y;
          if (px.is_zero() && py.is_zero()) continue;
          const uint32_t* lp = L + ((bk * GS_NUM_LINES + idx) * GS_LINE_WORDS) * nprob + p;
          // c0 straight to shared memory; c1 * xP and c2 * yP on the way
#pragma unroll
          for (int w = 0; w < 24; w++) lc[w * GS_MV2_NT] = lp[(size_t)w * nprob];
          fp t0, t1;
#pragma unroll
          for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int w = 0; w < 12; w++) {
              t0.l[w] = lp[(size_t)(24 + h * 12 + w) * nprob];
              t1.l[w] = lp[(size_t)(48 + h * 12 + w) * nprob];
            }
            fp::mul(t0, t0, px);
            fp::mul(t1, t1, py);
#pragma unroll
            for (int w = 0; w < 12; w++) {
              lc[(24 + h * 12 + w) * GS_MV2_NT] = t0.l[w];
              lc[(48 + h * 12 + w) * GS_MV2_NT] = t1.l[w];
            }
          }
          f12w_mul_line(f, lc, GS_MV2_NT);
        }
      }
    }
    f12w_store_conj(out, f, GS_MV2_NT);
  }
  F[((size_t)ch * 4 + e) * nprob + p] = out;
}

// ------------------------------------------------------------------ AoS -> slot scatter for ComT ops
// xs[p][k] (Com1), ys[p][k] (Com2) -> X, Y slot arrays
__global__ void k_scatter_pairs(const g1_aff* __restrict__ xs, const g2_aff* __restrict__ ys, g1_aff* __restrict__ X,
                                g2_aff* __restrict__ Y, size_t nprob, int K) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nprob * (size_t)K) return;
  size_t p = id % nprob;
  int k = (int)(id / nprob);
  for (int a = 0; a < 2; a++) {
    X[((size_t)a * K + k) * nprob + p] = xs[(p * K + k) * 2 + a];
    Y[((size_t)a * K + k) * nprob + p] = ys[(p * K + k) * 2 + a];
  }
}

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

}  // namespace gs

int gsi::pairing_init(gs_ctx* ctx) {
  CUDA_TRY(cudaFuncSetAttribute(k_miller, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_MV2_SMEM));
  return GS_OK;
}

// ------------------------------------------------------------------ pairing-product pipeline
// X, Y: device slot arrays [2][K][nprob].  Produces either ComT values (out_comt, AoS [p][4]) or
// per-entry verdict bytes ok4[4][nprob] (compared with 1 / target).
int gsi::run_pairing_product(gs_ctx* ctx, Scratch& sc, const g1_aff* X, const g2_aff* Y, size_t nprob, int K,
                               fp12* out_comt, uint8_t* ok4, const fp12* target) {
  size_t npoints = 2 * (size_t)K * nprob;
  uint32_t* L;
  uint8_t* yinf;
  fp12* F;
  CUDA_TRY(sc.alloc(&L, npoints * GS_NUM_LINES * GS_LINE_WORDS));
  CUDA_TRY(sc.alloc(&yinf, npoints));
  // split the slots over threads when there are few problems (one big statement)
  int S = K, nchunk = 1;
  size_t want_threads = 148 * 256;
  if (nprob * 4 < want_threads && K > 2) {
    size_t c = (want_threads + nprob * 4 - 1) / (nprob * 4);
    if (c > (size_t)(K + 1) / 2) c = (K + 1) / 2;  // at least 2 slots per chunk
    if (c < 1) c = 1;
    S = (int)((K + c - 1) / c);
    nchunk = (K + S - 1) / S;
  }
  CUDA_TRY(sc.alloc(&F, (size_t)nchunk * 4 * nprob));
  LAUNCH(k_g2_prepare, npoints, Y, L, yinf, npoints, nprob);
  {
    size_t nt_ = nprob * 4 * (size_t)nchunk;
    gs_ctx::prof_rec pr_{"k_miller", nullptr, nullptr};
    if (ctx->profile) {
      cudaEventCreate(&pr_.e0);
      cudaEventCreate(&pr_.e1);
      cudaEventRecord(pr_.e0, ctx->stream);
    }
    k_miller<<<(unsigned)((nt_ + GS_MV2_NT - 1) / GS_MV2_NT), GS_MV2_NT, GS_MV2_SMEM, ctx->stream>>>(X, yinf, L, F, nprob, K, S,
                                                                                                      nchunk);
    if (ctx->profile) {
      cudaEventRecord(pr_.e1, ctx->stream);
      ctx->prof.push_back(pr_);
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
  }
  return gsi::launch_final_exp(ctx, F, nprob, nchunk, out_comt, ok4, target);
}

static int comt_pairing_impl(gs_ctx* ctx, size_t nprob, int K, const gs_com1* xs, const gs_com2* ys, gs_comt* out) {
  if (!ctx || !xs || !ys || !out) return GS_EARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  g1_aff *dx, *X;
  g2_aff *dy, *Y;
  fp12* dout;
  size_t np = nprob * K;
  CUDA_TRY(upload(ctx, sc, &dx, xs, np * 2));
  CUDA_TRY(upload(ctx, sc, &dy, ys, np * 2));
  CUDA_TRY(sc.alloc(&X, np * 2));
  CUDA_TRY(sc.alloc(&Y, np * 2));
  CUDA_TRY(sc.alloc(&dout, nprob * 4));
  LAUNCH(k_scatter_pairs, np, dx, dy, X, Y, nprob, K);
  int rc = gsi::run_pairing_product(ctx, sc, X, Y, nprob, K, dout, nullptr, nullptr);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out, dout, nprob * 4 * sizeof(fp12), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}


extern "C" {

int gs_comt_pairing(gs_ctx* ctx, size_t count, const gs_com1* xs, const gs_com2* ys, gs_comt* out) {
  if (!ctx) return GS_EARG;
  if (count == 0) return GS_OK;
  return comt_pairing_impl(ctx, count, 1, xs, ys, out);
}

int gs_comt_pairing_sum(gs_ctx* ctx, size_t k, const gs_com1* xs, const gs_com2* ys, gs_comt* out) {
  if (!ctx || !out) return GS_EARG;
  if (k == 0) {  // empty sum = ComT::zero() = four GT identities
    CUDA_TRY(cudaSetDevice(ctx->device));
    Scratch sc(ctx);
    fp12* d;
    CUDA_TRY(sc.alloc(&d, 4));
    LAUNCH(k_fp12_set_one, 4, d, (size_t)4);
    CUDA_TRY(cudaMemcpyAsync(out, d, 4 * sizeof(fp12), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return GS_OK;
  }
  if (k > (size_t)1 << 24) FAIL(GS_EDIM, "pairing_sum: too many pairs");
  return comt_pairing_impl(ctx, 1, (int)k, xs, ys, out);
}

int gs_pairing(gs_ctx* ctx, size_t count, const gs_g1* ps, const gs_g2* qs, gs_gt* out) {
  if (!ctx || !ps || !qs || !out) return GS_EARG;
  if (count == 0) return GS_OK;
  // e(P,Q) = entry (0,0) of F((P,O),(Q,O))
  std::vector<gs_com1> xs(count);
  std::vector<gs_com2> ys(count);
  std::vector<gs_comt> res(count);
  memset(xs.data(), 0, count * sizeof(gs_com1));
  memset(ys.data(), 0, count * sizeof(gs_com2));
  for (size_t i = 0; i < count; i++) {
    xs[i].p[0] = ps[i];
    ys[i].p[0] = qs[i];
  }
  int rc = comt_pairing_impl(ctx, count, 1, xs.data(), ys.data(), res.data());
  if (rc) return rc;
  for (size_t i = 0; i < count; i++) out[i] = res[i].e[0];
  return GS_OK;
}

int gs_comt_linear_map(gs_ctx* ctx, int type, const void* target, gs_comt* out) {
  if (!ctx || !target || !out) return GS_EARG;
  if (type < 0 || type > 3) FAIL(GS_EARG, "linear_map: bad equation type");
  if (type != GS_PPE && !ctx->crs_loaded) FAIL(GS_EARG, "linear_map: no CRS loaded");
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  g1_aff* X;
  g2_aff* Y;
  fp12* dout;
  void* dt;
  size_t tsz = type == GS_PPE ? sizeof(fp12) : type == GS_MSMEG1 ? sizeof(g1_aff) : type == GS_MSMEG2 ? sizeof(g2_aff) : sizeof(fr);
  CUDA_TRY(upload(ctx, sc, (uint8_t**)&dt, target, tsz));
  CUDA_TRY(sc.alloc(&X, 2));
  CUDA_TRY(sc.alloc(&Y, 2));
  CUDA_TRY(sc.alloc(&dout, 4));
  if (type == GS_PPE) {
    LAUNCH(k_linear_map_ppe, 4, (const fp12*)dt, dout);
  } else {
    LAUNCH(k_linear_map_slots, 1, type, dt, ctx->crs, X, Y);
    int rc = gsi::run_pairing_product(ctx, sc, X, Y, 1, 1, dout, nullptr, nullptr);
    if (rc) return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(out, dout, 4 * sizeof(fp12), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}


}  // extern "C"
