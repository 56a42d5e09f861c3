// libgs_b200.so -- C ABI over the sm_100a kernels (include/gs_b200.h).
// No CPU fallback anywhere: every entry point launches CUDA kernels or fails with GS_ECUDA.
#include "../../include/gs_b200.h"

#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "prover_kernels.cuh"

using namespace gs;

static_assert(sizeof(gs_fr) == sizeof(fr), "fr layout");
static_assert(sizeof(gs_g1) == sizeof(g1_aff), "g1 layout");
static_assert(sizeof(gs_g2) == sizeof(g2_aff), "g2 layout");
static_assert(sizeof(gs_gt) == sizeof(fp12), "gt layout");
static_assert(sizeof(gs_com1) == 2 * sizeof(g1_aff), "com1 layout");
static_assert(sizeof(gs_com2) == 2 * sizeof(g2_aff), "com2 layout");
static_assert(sizeof(line_coeffs) == 288, "line layout");

struct gs_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  crs_dev* crs = nullptr;  // device
  bool crs_loaded = false;
  fixed_tables tabs;       // device fixed-base tables (prover_kernels.cuh)
  uint64_t launches = 0;
  bool profile = false;
  struct prof_rec {
    const char* name;
    cudaEvent_t e0, e1;
  };
  std::vector<prof_rec> prof;
  size_t verify_batch_max = 16384;  // problems per pass (bounds the line-coefficient scratch)
  std::string err;
};

#define CUDA_TRY(x)                                                                             \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      ctx->err = std::string(#x) + ": " + cudaGetErrorString(e_);                               \
      return GS_ECUDA;                                                                          \
    }                                                                                           \
  } while (0)
#define FAIL(code, msg)  \
  do {                   \
    ctx->err = (msg);    \
    return (code);       \
  } while (0)
#define LAUNCH(kern, nthreads, ...)                                                     \
  do {                                                                                  \
    size_t nt_ = (nthreads);                                                            \
    if (nt_ > 0) {                                                                      \
      unsigned grid_ = (unsigned)((nt_ + 127) / 128);                                   \
      gs_ctx::prof_rec pr_{#kern, nullptr, nullptr};                                    \
      if (ctx->profile) {                                                               \
        cudaEventCreate(&pr_.e0);                                                       \
        cudaEventCreate(&pr_.e1);                                                       \
        cudaEventRecord(pr_.e0, ctx->stream);                                           \
      }                                                                                 \
      kern<<<grid_, 128, 0, ctx->stream>>>(__VA_ARGS__);                                \
      if (ctx->profile) {                                                               \
        cudaEventRecord(pr_.e1, ctx->stream);                                           \
        ctx->prof.push_back(pr_);                                                       \
      }                                                                                 \
      ctx->launches++;                                                                  \
      CUDA_TRY(cudaGetLastError());                                                     \
    }                                                                                   \
  } while (0)

// stream-ordered scratch with RAII release
struct Scratch {
  gs_ctx* ctx;
  std::vector<void*> ptrs;
  explicit Scratch(gs_ctx* c) : ctx(c) {}
  template <class T>
  cudaError_t alloc(T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMallocAsync(&q, count * sizeof(T) + 16, ctx->stream);
    if (e == cudaSuccess) ptrs.push_back(q);
    *p = (T*)q;
    return e;
  }
  ~Scratch() {
    for (void* q : ptrs) cudaFreeAsync(q, ctx->stream);
  }
};

template <class T>
static cudaError_t upload(gs_ctx* ctx, Scratch& s, T** dst, const void* src, size_t count) {
  cudaError_t e = s.alloc(dst, count);
  if (e != cudaSuccess) return e;
  if (count == 0) return cudaSuccess;
  return cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
}

extern "C" {

int gs_ctx_create(int device, gs_ctx** out) {
  if (!out) return GS_EARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return GS_ECUDA;
  gs_ctx* ctx = new gs_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&ctx->crs, sizeof(crs_dev)) != cudaSuccess) {
    delete ctx;
    return GS_ECUDA;
  }
  // keep freed scratch in the pool instead of returning it to the driver after every call
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  // deep call chains (Fp12 -> Fp6 -> Fp2) with big local frames
  cudaDeviceSetLimit(cudaLimitStackSize, 32 * 1024);
  if (cudaFuncSetAttribute(k_miller, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_MV2_SMEM) != cudaSuccess) {
    delete ctx;
    return GS_ECUDA;
  }
  *out = ctx;
  return GS_OK;
}

void gs_ctx_destroy(gs_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->tabs.release();
  cudaFree(ctx->crs);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

const char* gs_last_error(const gs_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t gs_launch_count(const gs_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* gs_stream(const gs_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }


int gs_profile_enable(gs_ctx* ctx, int on) {
  if (!ctx) return GS_EARG;
  ctx->profile = on != 0;
  return GS_OK;
}

int gs_profile_read(gs_ctx* ctx, char* buf, size_t cap) {
  if (!ctx || !buf || cap == 0) return -GS_EARG;
  cudaSetDevice(ctx->device);
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return -GS_ECUDA;
  std::vector<std::string> names;
  std::vector<double> ms;
  std::vector<int> cnt;
  for (auto& r : ctx->prof) {
    float t = 0;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
    size_t i = 0;
    for (; i < names.size(); i++)
      if (names[i] == r.name) break;
    if (i == names.size()) {
      names.push_back(r.name);
      ms.push_back(0);
      cnt.push_back(0);
    }
    ms[i] += t;
    cnt[i]++;
  }
  ctx->prof.clear();
  std::string out;
  for (size_t i = 0; i < names.size(); i++) {
    char line[256];
    snprintf(line, sizeof line, "%s %d %.6f\n", names[i].c_str(), cnt[i], ms[i]);
    out += line;
  }
  size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
  memcpy(buf, out.data(), n);
  buf[n] = 0;
  return (int)n;
}

}  // extern "C"

// register-only Fp product chain (the measured integer-multiply roofline)
__global__ void __launch_bounds__(256) k_diag_fpmul(fp* out, const fp* in, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  fp x = in[t], y = in[t + 1];
  for (int i = 0; i < iters; i++) fp::mul(x, x, y);
  out[t] = x;
}

extern "C" int gs_diag_fpmul_rate(gs_ctx* ctx, double* fpmul_per_sec) {
  if (!ctx || !fpmul_per_sec) return GS_EARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, ctx->device));
  const int blocks = prop.multiProcessorCount * 2, threads = 256, iters = 4000;
  Scratch sc(ctx);
  fp *in, *out;
  CUDA_TRY(sc.alloc(&in, (size_t)blocks * threads + 1));
  CUDA_TRY(sc.alloc(&out, (size_t)blocks * threads));
  CUDA_TRY(cudaMemsetAsync(in, 1, ((size_t)blocks * threads + 1) * sizeof(fp), ctx->stream));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0, ctx->stream);
    k_diag_fpmul<<<blocks, threads, 0, ctx->stream>>>(out, in, iters);
    cudaEventRecord(e1, ctx->stream);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double rate = (double)blocks * threads * iters / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *fpmul_per_sec = best;
  return GS_OK;
}

// ------------------------------------------------------------------ pairing-product pipeline
// X, Y: device slot arrays [2][K][nprob].  Produces either ComT values (out_comt, AoS [p][4]) or
// per-entry verdict bytes ok4[4][nprob] (compared with 1 / target).
static int run_pairing_product(gs_ctx* ctx, Scratch& sc, const g1_aff* X, const g2_aff* Y, size_t nprob, int K,
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
  LAUNCH(k_final_exp, nprob * 4, F, nprob, nchunk, out_comt, ok4, target);
  return GS_OK;
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
  int rc = run_pairing_product(ctx, sc, X, Y, nprob, K, dout, nullptr, nullptr);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out, dout, nprob * 4 * sizeof(fp12), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

static int load_crs_device(gs_ctx* ctx, const gs_crs* crs) {
  crs_dev h;
  memset(&h, 0, sizeof(h));
  for (int k = 0; k < 2; k++)
    for (int a = 0; a < 2; a++) {
      memcpy(&h.u[k][a], &crs->u[k].p[a], sizeof(g1_aff));
      memcpy(&h.v[k][a], &crs->v[k].p[a], sizeof(g2_aff));
    }
  memcpy(&h.g1, &crs->g1_gen, sizeof(g1_aff));
  memcpy(&h.g2, &crs->g2_gen, sizeof(g2_aff));
  CUDA_TRY(cudaMemcpyAsync(ctx->crs, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  LAUNCH(k_crs_derive, 1, ctx->crs);
  int rc = ctx->tabs.build(ctx->stream, ctx->crs, &ctx->launches);
  if (rc != 0) FAIL(GS_ECUDA, "fixed-base table build failed");
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  ctx->crs_loaded = true;
  return GS_OK;
}

extern "C" {

int gs_crs_load(gs_ctx* ctx, const gs_crs* crs) {
  if (!ctx || !crs) return GS_EARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  return load_crs_device(ctx, crs);
}

int gs_crs_generate(gs_ctx* ctx, const gs_g1* p1, const gs_g2* p2, const gs_fr* a1, const gs_fr* a2, const gs_fr* t1,
                    const gs_fr* t2, gs_crs* out) {
  if (!ctx || !p1 || !p2 || !a1 || !a2 || !t1 || !t2 || !out) return GS_EARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  {
    Scratch sc(ctx);
    crs_gen_in hin;
    memcpy(&hin.p1, p1, sizeof(g1_aff));
    memcpy(&hin.p2, p2, sizeof(g2_aff));
    memcpy(&hin.a1, a1, sizeof(fr));
    memcpy(&hin.a2, a2, sizeof(fr));
    memcpy(&hin.t1, t1, sizeof(fr));
    memcpy(&hin.t2, t2, sizeof(fr));
    crs_gen_in* din;
    crs_gen_out* dout;
    CUDA_TRY(upload(ctx, sc, &din, &hin, 1));
    CUDA_TRY(sc.alloc(&dout, 1));
    LAUNCH(k_crs_generate, 6, din, dout);
    crs_gen_out hout;
    CUDA_TRY(cudaMemcpyAsync(&hout, dout, sizeof(hout), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    memset(out, 0, sizeof(*out));
    memcpy(&out->u[0].p[0], &hout.p1, sizeof(g1_aff));
    memcpy(&out->u[0].p[1], &hout.q1, sizeof(g1_aff));
    memcpy(&out->u[1].p[0], &hout.u1, sizeof(g1_aff));
    memcpy(&out->u[1].p[1], &hout.v1, sizeof(g1_aff));
    memcpy(&out->v[0].p[0], &hout.p2, sizeof(g2_aff));
    memcpy(&out->v[0].p[1], &hout.q2, sizeof(g2_aff));
    memcpy(&out->v[1].p[0], &hout.u2, sizeof(g2_aff));
    memcpy(&out->v[1].p[1], &hout.v2, sizeof(g2_aff));
    out->g1_gen = *p1;
    out->g2_gen = *p2;
  }
  int rc = gs_pairing(ctx, 1, p1, p2, &out->gt_gen);
  if (rc) return rc;
  return load_crs_device(ctx, out);
}

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
    int rc = run_pairing_product(ctx, sc, X, Y, 1, 1, dout, nullptr, nullptr);
    if (rc) return rc;
  }
  CUDA_TRY(cudaMemcpyAsync(out, dout, 4 * sizeof(fp12), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

// ------------------------------------------------------------------ verify
static size_t elem_size_A(int type) { return (type == 0 || type == 1) ? sizeof(g1_aff) : sizeof(fr); }
static size_t elem_size_B(int type) { return (type == 0 || type == 2) ? sizeof(g2_aff) : sizeof(fr); }
static size_t elem_size_T(int type) {
  return type == 0 ? sizeof(fp12) : type == 1 ? sizeof(g1_aff) : type == 2 ? sizeof(g2_aff) : sizeof(fr);
}

int gs_verify_batch_dev(gs_ctx* ctx, int type, size_t count, size_t m, size_t n, const void* a_consts,
                        const void* b_consts, const gs_fr* gamma, const void* target, const gs_com1* xcoms,
                        const gs_com2* ycoms, const gs_com2* pi, const gs_com1* theta, uint8_t* out_ok_dev) {
  if (!ctx) return GS_EARG;
  if (type < 0 || type > 3) FAIL(GS_EARG, "verify: bad equation type");
  if (!ctx->crs_loaded) FAIL(GS_EARG, "verify: no CRS loaded");
  if (count == 0) return GS_OK;
  // the reference panics on empty variable lists: keep it an error
  if (m == 0 || n == 0) FAIL(GS_EDIM, "verify: empty variable list");
  if (m > 1 << 20 || n > 1 << 20) FAIL(GS_EDIM, "verify: too many variables");
  if (!a_consts || !b_consts || !gamma || !target || !xcoms || !ycoms || !pi || !theta || !out_ok_dev) return GS_EARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  verify_shape s = make_verify_shape(type, (int)m, (int)n);
  for (size_t off = 0; off < count; off += ctx->verify_batch_max) {
    size_t nprob = count - off < ctx->verify_batch_max ? count - off : ctx->verify_batch_max;
    Scratch sc(ctx);
    verify_args v;
    v.a_consts = (const char*)a_consts + off * n * elem_size_A(type);
    v.b_consts = (const char*)b_consts + off * m * elem_size_B(type);
    v.gamma = (const fr*)gamma + off * m * n;
    v.target = (const char*)target + off * elem_size_T(type);
    v.xcoms = (const g1_aff*)xcoms + off * m * 2;
    v.ycoms = (const g2_aff*)ycoms + off * n * 2;
    v.pi = (const g2_aff*)pi + off * s.cx * 2;
    v.theta = (const g1_aff*)theta + off * s.cy * 2;
    g1_aff* X;
    g2_aff* Y;
    g1_jac* part;
    uint8_t* ok4;
    CUDA_TRY(sc.alloc(&X, 2 * (size_t)s.K * nprob));
    CUDA_TRY(sc.alloc(&Y, 2 * (size_t)s.K * nprob));
    CUDA_TRY(sc.alloc(&part, (size_t)s.nchunk * s.n_out * 2 * nprob));
    CUDA_TRY(sc.alloc(&ok4, 4 * nprob));
    LAUNCH(k_verify_assemble, nprob * (size_t)s.K, s, v, ctx->crs, X, Y, nprob);
    g1_aff* vtab;
    CUDA_TRY(sc.alloc(&vtab, (size_t)s.nbases * 2 * GS_VTAB * nprob));
    LAUNCH(k_vmsm_tables, nprob * (size_t)s.nbases * 2, s, v, ctx->crs, vtab, nprob);
    LAUNCH(k_vmsm_partial, nprob * (size_t)s.n_out * 2 * s.nchunk, s, v, vtab, part, nprob);
    LAUNCH(k_vmsm_reduce, nprob * (size_t)s.n_out * 2, s, v, part, X, nprob);
    int rc = run_pairing_product(ctx, sc, X, Y, nprob, s.K, nullptr, ok4, type == GS_PPE ? (const fp12*)v.target : nullptr);
    if (rc) return rc;
    LAUNCH(k_and4, nprob, ok4, out_ok_dev + off, nprob);
  }
  return GS_OK;
}

int gs_verify_batch(gs_ctx* ctx, int type, size_t count, size_t m, size_t n, const void* a_consts,
                    const void* b_consts, const gs_fr* gamma, const void* target, const gs_com1* xcoms,
                    const gs_com2* ycoms, const gs_com2* pi, const gs_com1* theta, uint8_t* out_ok) {
  if (!ctx) return GS_EARG;
  if (type < 0 || type > 3) FAIL(GS_EARG, "verify: bad equation type");
  if (count == 0) return GS_OK;
  if (m == 0 || n == 0) FAIL(GS_EDIM, "verify: empty variable list");
  if (!a_consts || !b_consts || !gamma || !target || !xcoms || !ycoms || !pi || !theta || !out_ok) return GS_EARG;
  CUDA_TRY(cudaSetDevice(ctx->device));
  verify_shape s = make_verify_shape(type, (int)m, (int)n);
  Scratch sc(ctx);
  uint8_t *dA, *dB, *dT, *dok;
  fr* dG;
  g1_aff *dc, *dth;
  g2_aff *dd, *dpi;
  CUDA_TRY(upload(ctx, sc, &dA, a_consts, count * n * elem_size_A(type)));
  CUDA_TRY(upload(ctx, sc, &dB, b_consts, count * m * elem_size_B(type)));
  CUDA_TRY(upload(ctx, sc, &dG, gamma, count * m * n));
  CUDA_TRY(upload(ctx, sc, &dT, target, count * elem_size_T(type)));
  CUDA_TRY(upload(ctx, sc, &dc, xcoms, count * m * 2));
  CUDA_TRY(upload(ctx, sc, &dd, ycoms, count * n * 2));
  CUDA_TRY(upload(ctx, sc, &dpi, pi, count * s.cx * 2));
  CUDA_TRY(upload(ctx, sc, &dth, theta, count * s.cy * 2));
  CUDA_TRY(sc.alloc(&dok, count));
  int rc = gs_verify_batch_dev(ctx, type, count, m, n, dA, dB, (const gs_fr*)dG, dT, (const gs_com1*)dc,
                               (const gs_com2*)dd, (const gs_com2*)dpi, (const gs_com1*)dth, dok);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out_ok, dok, count, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

}  // extern "C"

#include "prover_abi.inc"
