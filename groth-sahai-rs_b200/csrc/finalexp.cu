// Final exponentiation kernel (arkworks exponent: easy part, then (x-1)^2 (x+p)(x^2+p^2-1) + 3) and the
// comparison against the expected ComT entry.  Reference: ark-ec final_exponentiation as reached from
// src/data_structures.rs:484-502.
#include "ctx.h"
#include "coop12.cuh"
#include "pairing.cuh"

using namespace gs;

namespace gs {

// ------------------------------------------------------------------ cooperative final exponentiation (+ compare)
// block = 32 instances id = p + e*nprob (6 warps, one per w-power coefficient, coop12.cuh); the op program of
// cq_build_final_exp runs over 5 accumulator buffers in shared memory with one barrier per op.
constexpr int FE3_SMEM = CQ_FE_NBUF * CQ_ACC * 4 + CQ_LANES * 4;  // per 6-warp group
__global__ void __launch_bounds__(CQ_BLOCK_THREADS, 1) k_final_exp3(const fp12* __restrict__ F, size_t nprob, int nchunk,
                                                                   fp12* __restrict__ out_comt, uint8_t* __restrict__ ok,
                                                                   const fp12* __restrict__ target,
                                                                   const uint32_t* __restrict__ prog, int nops, size_t ngroups, int ne) {
  extern __shared__ __align__(16) uint32_t sm_all[];
  const int grp = threadIdx.x / CQ_GROUP_THREADS, tg = threadIdx.x % CQ_GROUP_THREADS;
  uint32_t* bufs = sm_all + (size_t)grp * (FE3_SMEM / 4);
  uint32_t* bad = bufs + CQ_FE_NBUF * CQ_ACC;
  const int k = tg >> 5, lane = tg & 31;
  const size_t gid = (size_t)blockIdx.x * CQ_GROUPS + grp;
  if (gid >= ngroups) return;
  const size_t id = gid * CQ_LANES + lane;
  const bool valid = id < nprob * ne;
  const size_t p = valid ? id % nprob : 0;
  const int e = valid ? (int)(id / nprob) : 0;
  const int pos = cq_tower_pos(k);
  if (k == 0) bad[lane] = 0;
  // f = product of the chunk partial products
  for (int ch = 0; ch < nchunk; ch++) {
    fp2 c;
    if (valid) {
      c = ((const fp2*)&F[((size_t)ch * ne + e) * nprob + p])[pos];
    } else {
      c.set_zero();
      if (k == 0) fp_one(c.c0);
    }
    cq_st_coef(bufs + (ch == 0 ? 0 : 1) * CQ_ACC, k, lane, c);
    cq_group_sync(grp);
    if (ch > 0) {
      cq_mul(k, lane, bufs, bufs + CQ_ACC, bufs + 2 * CQ_ACC);
      cq_group_sync(grp);
      cq_copy(k, lane, bufs + 2 * CQ_ACC, bufs);
      cq_group_sync(grp);
    }
  }
#pragma unroll 1
  for (int i = 0; i < nops; i++) {
    cq_exec(prog[i], k, lane, bufs);
    cq_group_sync(grp);
  }
  fp2 g;
  cq_ld_coef(g.c0, g.c1, bufs + CQ_FE_OUT * CQ_ACC, k, lane, false, false);
  if (valid) {
    if (out_comt) ((fp2*)&out_comt[p * ne + e])[pos] = g;
    if (ok) {
      fp2 want;
      if (target != nullptr && e == ne - 1) {
        want = ((const fp2*)&target[p])[pos];
      } else {
        want.set_zero();
        if (k == 0) fp_one(want.c0);
      }
      if (!g.equals(want)) bad[lane] = 1;
    }
  }
  cq_group_sync(grp);
  if (ok && valid && k == 0) ok[(size_t)e * nprob + p] = bad[lane] ? 0 : 1;
}

// ------------------------------------------------------------------ chunk-product tree (big statements)
// A statement with few problems and many slots is split into up to ~1,000 accumulator chunks (pairing.cu);
// multiplying them one after the other inside k_final_exp3 cost 23 ms of a 137 ms verify at m = n = 1024.
// Here 32 lanes x 6 warps multiply L chunks each:  F2[(part*4 + e)*nprob + p] = prod_{ch in part} F[(ch*4 + e)*nprob + p]
constexpr int CR_SMEM = 3 * CQ_ACC * 4;  // per 6-warp group
__global__ void __launch_bounds__(CQ_BLOCK_THREADS, 1) k_chunk_reduce(const fp12* __restrict__ F, fp12* __restrict__ F2, size_t nprob,
                                                                     int nchunk, int L, int nparts, size_t ngroups, int ne) {
  extern __shared__ __align__(16) uint32_t sm_all[];
  const int grp = threadIdx.x / CQ_GROUP_THREADS, tg = threadIdx.x % CQ_GROUP_THREADS;
  uint32_t* bufs = sm_all + (size_t)grp * (CR_SMEM / 4);
  const int k = tg >> 5, lane = tg & 31;
  const size_t gid = (size_t)blockIdx.x * CQ_GROUPS + grp;
  if (gid >= ngroups) return;
  const size_t id = gid * CQ_LANES + lane;
  const bool valid = id < nprob * ne * (size_t)nparts;
  const size_t p = valid ? id % nprob : 0;
  const int e = valid ? (int)((id / nprob) % ne) : 0;
  const int part = valid ? (int)(id / (nprob * ne)) : 0;
  const int pos = cq_tower_pos(k);
  int cur = 0;
  for (int i = 0; i < L; i++) {
    const int ch = part * L + i;
    fp2 c;
    if (valid && ch < nchunk) {
      c = ((const fp2*)&F[((size_t)ch * ne + e) * nprob + p])[pos];
    } else {
      c.set_zero();
      if (k == 0) fp_one(c.c0);
    }
    if (i == 0) {
      cq_st_coef(bufs, k, lane, c);
      cq_group_sync(grp);
    } else {
      cq_st_coef(bufs + 2 * CQ_ACC, k, lane, c);
      cq_group_sync(grp);
      cq_mul(k, lane, bufs + cur * CQ_ACC, bufs + 2 * CQ_ACC, bufs + (cur ^ 1) * CQ_ACC);
      cq_group_sync(grp);
      cur ^= 1;
    }
  }
  fp2 g;
  cq_ld_coef(g.c0, g.c1, bufs + cur * CQ_ACC, k, lane, false, false);
  if (valid) ((fp2*)&F2[((size_t)part * ne + e) * nprob + p])[pos] = g;
}

// ------------------------------------------------------------------ GT powers with 64-bit exponents (gs_verify_batch_rand)
// lane = one element; out = t^e.  The multiplication of a window picks ITS lane's table entry (the X operand of cq_mul is
// streamed from shared memory, so a per-lane buffer is only a per-lane base address); lanes whose digit is 0 copy through.
__global__ void __launch_bounds__(CQ_BLOCK_THREADS, 1) k_gt_pow64(const fp12* __restrict__ T, const uint64_t* __restrict__ E, size_t estride,
                                                                 size_t count, fp12* __restrict__ out, size_t ngroups) {
  extern __shared__ __align__(16) uint32_t sm_all[];
  const int grp = threadIdx.x / CQ_GROUP_THREADS, tg = threadIdx.x % CQ_GROUP_THREADS;
  uint32_t* bufs = sm_all + (size_t)grp * (FE3_SMEM / 4);
  const int k = tg >> 5, lane = tg & 31;
  const size_t gid = (size_t)blockIdx.x * CQ_GROUPS + grp;
  if (gid >= ngroups) return;
  const size_t id = gid * CQ_LANES + lane;
  const bool valid = id < count;
  const int pos = cq_tower_pos(k);
  const uint64_t ex = valid ? E[id * estride] : 0;
  fp2 c;
  if (valid) {
    c = ((const fp2*)&T[id])[pos];
  } else {
    c.set_zero();
    if (k == 0) fp_one(c.c0);
  }
  cq_st_coef(bufs + 2 * CQ_ACC, k, lane, c);  // t
  cq_group_sync(grp);
  cq_cyc_sqr(k, lane, bufs + 2 * CQ_ACC, bufs + 3 * CQ_ACC);  // t^2
  cq_group_sync(grp);
  cq_mul(k, lane, bufs + 2 * CQ_ACC, bufs + 3 * CQ_ACC, bufs + 4 * CQ_ACC);  // t^3
  cq_group_sync(grp);
  {  // top window
    const int d = (int)(ex >> 62);
    if (d == 0)
      cq_set_one(k, lane, bufs);
    else
      cq_copy(k, lane, bufs + (1 + d) * CQ_ACC, bufs);
  }
  cq_group_sync(grp);
  int cur = 0;
#pragma unroll 1
  for (int w = 30; w >= 0; w--) {
    cq_cyc_sqr(k, lane, bufs + cur * CQ_ACC, bufs + (cur ^ 1) * CQ_ACC);
    cq_group_sync(grp);
    cq_cyc_sqr(k, lane, bufs + (cur ^ 1) * CQ_ACC, bufs + cur * CQ_ACC);
    cq_group_sync(grp);
    const int d = (int)((ex >> (2 * w)) & 3);
    cq_mul(k, lane, bufs + cur * CQ_ACC, bufs + (1 + (d ? d : 1)) * CQ_ACC, bufs + (cur ^ 1) * CQ_ACC);
    if (d == 0) cq_copy(k, lane, bufs + cur * CQ_ACC, bufs + (cur ^ 1) * CQ_ACC);
    cq_group_sync(grp);
    cur ^= 1;
  }
  fp2 g;
  cq_ld_coef(g.c0, g.c1, bufs + cur * CQ_ACC, k, lane, false, false);
  if (valid) ((fp2*)&out[id])[pos] = g;
}

}  // namespace gs

// Reduces the chunk dimension of F ([nchunk][4][nprob]) until at most `max_out` chunks are left; *F then points
// at scratch owned by `sc`.
int gsi::reduce_chunks(gs_ctx* ctx, Scratch& sc, const fp12** F, size_t nprob, int* nchunk, int max_out, int ne) {
  if (max_out < 1) max_out = 1;
  while (*nchunk > max_out) {
    int L = 2;
    while (L * L < *nchunk) L++;  // ~sqrt: serial depth of this pass ~ serial depth left for the consumer
    if (L > 64) L = 64;
    if (ne == 1 && (size_t)*nchunk * nprob > 8192) {  // a long product (tens of thousands of factors): keep ~4,096 lanes busy per level
      L = (int)((size_t)*nchunk * nprob / 4096);
      if (L > 64) L = 64;
    }
    const int nparts = (*nchunk + L - 1) / L;
    fp12* F2;
    CUDA_TRY(sc.alloc(&F2, (size_t)nparts * ne * nprob));
    size_t ngroups = (nprob * ne * (size_t)nparts + CQ_LANES - 1) / CQ_LANES;
    LAUNCH_CFG(k_chunk_reduce, ((ngroups + CQ_GROUPS - 1) / CQ_GROUPS) * CQ_BLOCK_THREADS, CQ_BLOCK_THREADS, CQ_GROUPS * CR_SMEM, *F,
               F2, nprob, *nchunk, L, nparts, ngroups, ne);
    *F = F2;
    *nchunk = nparts;
  }
  return GS_OK;
}

int gsi::final_exp_init(gs_ctx* ctx) {
  static uint32_t prog[CQ_FE_MAXOPS];
  int n = cq_build_final_exp(prog);
  ctx->fe_nops = n;
  CUDA_TRY(cudaMalloc(&ctx->fe_prog, n * sizeof(uint32_t)));
  CUDA_TRY(cudaMemcpy(ctx->fe_prog, prog, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaFuncSetAttribute(k_final_exp3, cudaFuncAttributeMaxDynamicSharedMemorySize, CQ_GROUPS * FE3_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(k_chunk_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, CQ_GROUPS * CR_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(k_gt_pow64, cudaFuncAttributeMaxDynamicSharedMemorySize, CQ_GROUPS * FE3_SMEM));
  return GS_OK;
}

int gsi::launch_final_exp(gs_ctx* ctx, const fp12* F, size_t nprob, int nchunk, fp12* out_comt, uint8_t* ok4, const fp12* target, int ne) {
  size_t ngroups = (nprob * ne + CQ_LANES - 1) / CQ_LANES;
  LAUNCH_CFG(k_final_exp3, ((ngroups + CQ_GROUPS - 1) / CQ_GROUPS) * CQ_BLOCK_THREADS, CQ_BLOCK_THREADS, CQ_GROUPS * FE3_SMEM, F, nprob,
             nchunk, out_comt, ok4, target, ctx->fe_prog, ctx->fe_nops, ngroups, ne);
  return GS_OK;
}

// t^e for 64-bit exponents, 32 proofs per 6-warp group: 2-bit windows over the table t, t^2, t^3 (buffers 2, 3, 4), the
// running power ping-pongs between buffers 0 and 1; squarings are Granger-Scott (t must be in the cyclotomic subgroup).
int gsi::gt_pow64(gs_ctx* ctx, const fp12* t, const uint64_t* e, size_t estride, size_t count, fp12* out) {
  size_t ngroups = (count + CQ_LANES - 1) / CQ_LANES;
  LAUNCH_CFG(k_gt_pow64, ((ngroups + CQ_GROUPS - 1) / CQ_GROUPS) * CQ_BLOCK_THREADS, CQ_BLOCK_THREADS, CQ_GROUPS * FE3_SMEM, t, e, estride,
             count, out, ngroups);
  return GS_OK;
}
