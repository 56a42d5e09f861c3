// Pairing-product pipeline: G2 line preparation, Miller accumulation, final exponentiation, and the
// ComT entry points of the C ABI (ComT::pairing / pairing_sum / linear_map_*, E::pairing).
// Reference: src/data_structures.rs:484-540, src/generator.rs:116.
#include "ctx.h"
#include "coop12.cuh"
#include "pairing.cuh"

using namespace gs;

namespace gs {

// "Slot" layout used by every pairing-product evaluation (ComT::pairing, pairing_sum, verify): a problem
// is K pairs (X_k in Com1, Y_k in Com2), and the wanted ComT is
//     ComT[a][b] = prod_k e(X_k.a, Y_k.b)          (src/data_structures.rs:494-502)
// Points are stored SoA over problems so that a warp (32 consecutive problems, same slot, same
// coordinate) reads contiguous memory:
//     X[(a*K + k) * nprob + p]   g1_aff        Y[(b*K + k) * nprob + p]   g2_aff

// ------------------------------------------------------------------ v3: evaluated line tiles + cooperative Miller
// Accumulator index A = chunk * np + (p - p0) (chunk = slot range [chunk*S, chunk*S+S) of a big statement);
// 32 consecutive accumulators of one ComT entry e = 2a+b form a block  bid = (A/32)*4 + e.
// Tile (bid, kk, step): the line of slot kk at Miller step `step`, already evaluated at the G1 point of
// entry e, as 6 Fp in the Q layout of coop12.cuh (alpha.c0 alpha.c1 beta.c0 beta.c1 gamma.c0 gamma.c1 =
// c0, c1*xP, c2*yP), 9,216 B, contiguous: written by k_g2_prepare3 with 16-B stores (512 B per warp and
// quad), copied into shared memory by k_miller3 with cp.async.
//     tiles[((bid*S + kk)*68 + step) * M3_TILE ...]      masks[bid*S + kk] = lanes whose pair is not dropped
constexpr int M3_NV = 6;
constexpr int M3_TILE = M3_NV * CQ_FP;
constexpr int M3_THREADS = 6 * CQ_LANES;
constexpr int M3_MAXS = 512;
constexpr int M3_SMEM = (2 * CQ_ACC + 2 * M3_TILE) * 4 + M3_MAXS * 6 + 16;

// one thread per G2 point q = (b*K + k) * np + pl  (pl = p - p0)
__global__ void __launch_bounds__(128) k_g2_prepare3(const g1_aff* __restrict__ X, const g2_aff* __restrict__ Y,
                                                     uint32_t* __restrict__ tiles, uint32_t* __restrict__ masks,
                                                     size_t nprob, size_t p0, size_t np, int K, int S) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= 2 * (size_t)K * np) return;
  size_t pl = q % np, bk = q / np;
  int b = (int)(bk / K), k = (int)(bk % K);
  g2_aff Q = Y[bk * nprob + p0 + pl];
  if (Q.is_inf()) return;
  fp px[2], py[2];
  bool act[2];
  size_t tbase[2];
  const int ch = k / S, kk = k % S;
  const size_t A = (size_t)ch * np + pl;
  const int lane = (int)(A & 31);
#pragma unroll
  for (int a = 0; a < 2; a++) {
    const g1_aff* P = &X[((size_t)a * K + k) * nprob + p0 + pl];
    px[a] = P->x;
    py[a] = P->y;
    act[a] = !(px[a].is_zero() && py[a].is_zero());
    size_t bid = (A >> 5) * 4 + (size_t)(2 * a + b);
    tbase[a] = ((bid * S + kk) * GS_NUM_LINES) * (size_t)M3_TILE;
    if (act[a]) atomicOr(&masks[bid * S + kk], 1u << lane);
  }
  if (!act[0] && !act[1]) return;
  g2_proj t;
  t.x = Q.x;
  t.y = Q.y;
  t.z.set_one();
  int idx = 0;
  for (int bit = 62; bit >= 0; bit--) {
    int nl = ((GS_X_ABS >> bit) & 1) ? 2 : 1;
    for (int w = 0; w < nl; w++, idx++) {
      line_coeffs l;
      if (w == 0)
        g2_double_step(t, l);
      else
        g2_add_step(t, Q, l);
#pragma unroll 1
      for (int a = 0; a < 2; a++) {
        if (!act[a]) continue;
        uint32_t* o = tiles + tbase[a] + (size_t)idx * M3_TILE;
        fp v;
        cq_st(cq_ptr(o, 0, lane), l.c0.c0);
        cq_st(cq_ptr(o, 1, lane), l.c0.c1);
        fp::mul(v, l.c1.c0, px[a]);
        cq_st(cq_ptr(o, 2, lane), v);
        fp::mul(v, l.c1.c1, px[a]);
        cq_st(cq_ptr(o, 3, lane), v);
        fp::mul(v, l.c2.c0, py[a]);
        cq_st(cq_ptr(o, 4, lane), v);
        fp::mul(v, l.c2.c1, py[a]);
        cq_st(cq_ptr(o, 5, lane), v);
      }
    }
  }
}

__device__ GS_INL void cp_async16(uint32_t* smem, const uint32_t* g) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
__device__ GS_INL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// block = 32 accumulators of entry e (bid = blk*4 + e); warp k = w-power coefficient k (coop12.cuh).
// F[(ch*4 + e) * nprob + p] = conj( prod over the chunk's slots )
__global__ void __launch_bounds__(M3_THREADS, 2) k_miller3(const uint32_t* __restrict__ tiles, const uint32_t* __restrict__ masks,
                                                          fp12* __restrict__ F, size_t nprob, size_t p0, size_t np, int S,
                                                          int nchunk) {
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* acc = sm;
  uint32_t* tile = sm + 2 * CQ_ACC;
  uint32_t* smask = tile + 2 * M3_TILE;
  uint16_t* slots = (uint16_t*)(smask + M3_MAXS);
  int* nact_s = (int*)(slots + M3_MAXS);
  const int k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t bid = blockIdx.x;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int kk = 0; kk < S; kk++) {
      uint32_t m = masks[bid * S + kk];
      if (m) {
        slots[n] = (uint16_t)kk;
        smask[n] = m;
        n++;
      }
    }
    *nact_s = n;
  }
  cq_set_one(k, lane, acc);
  __syncthreads();
  const int nact = *nact_s;
  int cur = 0;
  if (nact > 0) {
    const int total = GS_NUM_LINES * nact;
    auto issue = [&](int n, int stage) {
      int s = n / nact, i = n - s * nact;
      const uint32_t* src = tiles + ((bid * S + slots[i]) * GS_NUM_LINES + s) * (size_t)M3_TILE;
      uint32_t* dst = tile + stage * M3_TILE;
#pragma unroll
      for (int c = 0; c < M3_TILE / 4 / M3_THREADS; c++) {
        int w = (c * M3_THREADS + threadIdx.x) * 4;
        cp_async16(dst + w, src + w);
      }
    };
    issue(0, 0);
    cp_async_wait_all();
    __syncthreads();
    int n = 0;
    for (int bit = 62; bit >= 0; bit--) {
      if (bit != 62) {
        cq_sqr(k, lane, acc + cur * CQ_ACC, acc + (cur ^ 1) * CQ_ACC);
        __syncthreads();
        cur ^= 1;
      }
      int nl = ((GS_X_ABS >> bit) & 1) ? 2 : 1;
      for (int i = 0; i < nl * nact; i++, n++) {
        if (n + 1 < total) issue(n + 1, (n + 1) & 1);
        int si = i >= nact ? i - nact : i;
        bool active = (smask[si] >> lane) & 1;
        cq_line_mul(k, lane, acc + cur * CQ_ACC, acc + (cur ^ 1) * CQ_ACC, tile + (n & 1) * M3_TILE, active);
        cp_async_wait_all();
        __syncthreads();
        cur ^= 1;
      }
    }
  }
  // conjugate (x < 0) and write out in tower order
  size_t A = (bid >> 2) * 32 + lane;
  int e = (int)(bid & 3);
  if (A < np * (size_t)nchunk) {
    size_t pl = A % np;
    int ch = (int)(A / np);
    fp2 r;
    cq_ld_coef(r.c0, r.c1, acc + cur * CQ_ACC, k, lane, false, false);
    if (k & 1) fp2::neg(r, r);
    fp2* dst = (fp2*)&F[((size_t)ch * 4 + e) * nprob + p0 + pl];
    dst[cq_tower_pos(k)] = r;
  }
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
  CUDA_TRY(cudaFuncSetAttribute(k_miller3, cudaFuncAttributeMaxDynamicSharedMemorySize, M3_SMEM));
  return GS_OK;
}

// ------------------------------------------------------------------ pairing-product pipeline
// X, Y: device slot arrays [2][K][nprob].  Produces either ComT values (out_comt, AoS [p][4]) or
// per-entry verdict bytes ok4[4][nprob] (compared with 1 / target).
// Problems are processed in passes of `pc` so that the evaluated-line tiles stay within ctx->tile_budget
// bytes of HBM; a pass is sized to a whole number of k_miller3 waves (2 blocks x 148 SMs x 32 accumulators
// / 4 entries = 2,368 problems per wave) when the batch is large enough.
int gsi::run_pairing_product(gs_ctx* ctx, Scratch& sc, const g1_aff* X, const g2_aff* Y, size_t nprob, int K,
                               fp12* out_comt, uint8_t* ok4, const fp12* target) {
  const size_t wave = 2368;
  // split the slots of a big statement over several accumulators when there are few problems
  int S = K, nchunk = 1;
  if (nprob < wave && K > 2) {
    size_t c = (wave + nprob - 1) / nprob;
    if (c > (size_t)(K + 1) / 2) c = (K + 1) / 2;  // at least 2 slots per chunk
    if (c < 1) c = 1;
    S = (int)((K + c - 1) / c);
  }
  if (S > M3_MAXS) S = M3_MAXS;
  nchunk = (K + S - 1) / S;
  const size_t per_prob = (size_t)4 * nchunk * S * GS_NUM_LINES * M3_TILE * 4 / 32;  // tile bytes per problem
  size_t pc = ctx->tile_budget / per_prob;
  if (pc >= nprob) {
    pc = nprob;
  } else {
    if (pc > wave) pc -= pc % wave;
    if (pc < 32) pc = 32;
    pc -= pc % 32;
  }
  const size_t nblk_max = ((pc * nchunk + 31) / 32) * 4;
  uint32_t *tiles, *masks;
  fp12* F;
  CUDA_TRY(sc.alloc(&tiles, nblk_max * S * GS_NUM_LINES * (size_t)M3_TILE));
  CUDA_TRY(sc.alloc(&masks, nblk_max * S));
  CUDA_TRY(sc.alloc(&F, (size_t)nchunk * 4 * nprob));
  for (size_t p0 = 0; p0 < nprob; p0 += pc) {
    size_t np = nprob - p0 < pc ? nprob - p0 : pc;
    size_t nblk = ((np * nchunk + 31) / 32) * 4;
    CUDA_TRY(cudaMemsetAsync(masks, 0, nblk * S * sizeof(uint32_t), ctx->stream));
    LAUNCH(k_g2_prepare3, 2 * (size_t)K * np, X, Y, tiles, masks, nprob, p0, np, K, S);
    LAUNCH_CFG(k_miller3, nblk * M3_THREADS, M3_THREADS, M3_SMEM, tiles, masks, F, nprob, p0, np, S, nchunk);
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
