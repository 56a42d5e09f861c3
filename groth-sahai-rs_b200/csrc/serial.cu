// Wire formats either side of the hot path: batched point (de)compression with full
// validation, and canonical scalar / GT bytes.  Replaces ark-serialize's CanonicalSerialize/Deserialize as the
// reference derives it for CRS (generator.rs:35), Commit1/Commit2 (prover/commit.rs:18-28), EquProof
// (prover/prove.rs:55-61) and the equations (statement.rs:117-185); the point encoding is ark-bls12-381's
// zcash / IETF format: big-endian x (G2: x.c1 || x.c0), flag bits 0x80 compressed, 0x40 infinity, 0x20 y is the
// lexicographically largest of {y, -y}.  A service that verifies 65,536 proofs first has to decompress and
// subgroup-check ~1.8 M points: seconds of host time per batch, which is why this lives on the GPU.
// One thread per point: x -> y by a field square root, sign by the flag, membership by the endomorphism tests
// arkworks itself uses (two 64-bit scalar multiplications instead of [r]P = O).
#include "ctx.h"
#include "wire.cuh"

using namespace gs;

namespace gs {

// ------------------------------------------------------------------ compressed points: one thread per point (wire.cuh)
__global__ void __launch_bounds__(128) k_g1_compress(const g1_aff* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[48];
  g1_compress_point(b, in[i]);
  for (int j = 0; j < 48; j++) out[i * 48 + j] = b[j];
}
__global__ void __launch_bounds__(128) k_g1_decompress(const uint8_t* __restrict__ in, g1_aff* __restrict__ out, uint8_t* __restrict__ ok,
                                                       size_t n, int check_subgroup) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[48];
  for (int j = 0; j < 48; j++) b[j] = in[i * 48 + j];
  g1_aff p;
  const bool good = g1_decompress_point(p, b, check_subgroup);
  out[i] = p;
  ok[i] = good ? 1 : 0;
}
__global__ void __launch_bounds__(128) k_g2_compress(const g2_aff* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[96];
  g2_compress_point(b, in[i]);
  for (int j = 0; j < 96; j++) out[i * 96 + j] = b[j];
}
__global__ void __launch_bounds__(128) k_g2_decompress(const uint8_t* __restrict__ in, g2_aff* __restrict__ out, uint8_t* __restrict__ ok,
                                                       size_t n, int check_subgroup) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[96];
  for (int j = 0; j < 96; j++) b[j] = in[i * 96 + j];
  g2_aff p;
  const bool good = g2_decompress_point(p, b, check_subgroup);
  out[i] = p;
  ok[i] = good ? 1 : 0;
}

// ------------------------------------------------------------------ uncompressed encodings (serialize_uncompressed)
// G1: 96 B = x || y big-endian, G2: 192 B = x.c1 || x.c0 || y.c1 || y.c0; flag bits of byte 0: 0x80 must be clear,
// 0x40 = infinity (0x20, the sort flag, is not used).  Reading validates: coordinates < p, on the curve, in the subgroup.
__global__ void __launch_bounds__(128) k_g1_uncompressed(const g1_aff* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  g1_aff p = in[i];
  uint8_t b[96];
  for (int j = 0; j < 96; j++) b[j] = 0;
  if (p.is_inf()) {
    b[0] = 0x40;
  } else {
    fp_to_be(b, p.x);
    fp_to_be(b + 48, p.y);
  }
  for (int j = 0; j < 96; j++) out[i * 96 + j] = b[j];
}
__global__ void __launch_bounds__(128) k_g1_from_uncompressed(const uint8_t* __restrict__ in, g1_aff* __restrict__ out,
                                                              uint8_t* __restrict__ ok, size_t n, int check_subgroup) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[96];
  for (int j = 0; j < 96; j++) b[j] = in[i * 96 + j];
  g1_aff p;
  p.set_inf();
  bool good = (b[0] & 0x80) == 0;
  if (good && (b[0] & 0x40)) good = wire_inf_canonical(b, 96);
  if (good && !(b[0] & 0x40)) {
    good = (b[0] & 0x20) == 0 && (b[48] & 0xE0) == 0 && fp_from_be(p.x, b) && fp_from_be(p.y, b + 48);
    if (good) {
      fp lhs, rhs, four;
      fp::sqr(lhs, p.y);
      fp::sqr(rhs, p.x);
      fp::mul(rhs, rhs, p.x);
      for (int j = 0; j < 12; j++) four.l[j] = FP_FOUR(j);
      fp::add(rhs, rhs, four);
      good = lhs.equals(rhs);
      if (good && check_subgroup) good = in_subgroup_g1(p);
    }
    if (!good) p.set_inf();
  }
  out[i] = p;
  ok[i] = good ? 1 : 0;
}
__global__ void __launch_bounds__(128) k_g2_uncompressed(const g2_aff* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  g2_aff p = in[i];
  uint8_t b[192];
  for (int j = 0; j < 192; j++) b[j] = 0;
  if (p.is_inf()) {
    b[0] = 0x40;
  } else {
    fp_to_be(b, p.x.c1);
    fp_to_be(b + 48, p.x.c0);
    fp_to_be(b + 96, p.y.c1);
    fp_to_be(b + 144, p.y.c0);
  }
  for (int j = 0; j < 192; j++) out[i * 192 + j] = b[j];
}
__global__ void __launch_bounds__(128) k_g2_from_uncompressed(const uint8_t* __restrict__ in, g2_aff* __restrict__ out,
                                                              uint8_t* __restrict__ ok, size_t n, int check_subgroup) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[192];
  for (int j = 0; j < 192; j++) b[j] = in[i * 192 + j];
  g2_aff p;
  p.set_inf();
  bool good = (b[0] & 0x80) == 0;
  if (good && (b[0] & 0x40)) good = wire_inf_canonical(b, 192);
  if (good && !(b[0] & 0x40)) {
    good = (b[0] & 0x20) == 0 && ((b[48] | b[96] | b[144]) & 0xE0) == 0 && fp_from_be(p.x.c1, b) && fp_from_be(p.x.c0, b + 48) &&
           fp_from_be(p.y.c1, b + 96) && fp_from_be(p.y.c0, b + 144);
    if (good) {
      fp2 lhs, rhs, bt;
      fp2::sqr(lhs, p.y);
      fp2::sqr(rhs, p.x);
      fp2::mul(rhs, rhs, p.x);
      for (int j = 0; j < 12; j++) bt.c0.l[j] = bt.c1.l[j] = FP_FOUR(j);
      fp2::add(rhs, rhs, bt);
      good = lhs.equals(rhs);
      if (good && check_subgroup) good = in_subgroup_g2(p);
    }
    if (!good) p.set_inf();
  }
  out[i] = p;
  ok[i] = good ? 1 : 0;
}

// ------------------------------------------------------------------ scalars and GT: canonical little-endian integers
// Fr: 32 B LE (ark-ff CanonicalSerialize for Fp<4 limbs>); out-of-range input is rejected
__global__ void k_fr_to_bytes(const fr* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[8];
  fr_from_mont(k, in[i]);
  for (int j = 0; j < 8; j++) out[i * 8 + j] = k[j];
}
__global__ void k_fr_from_bytes(const uint32_t* __restrict__ in, fr* __restrict__ out, uint8_t* __restrict__ ok, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t l[8], m[8];
  fr raw, r2, r;
  for (int j = 0; j < 8; j++) {
    l[j] = in[i * 8 + j];
    m[j] = FrParams::mod(j);
    raw.l[j] = l[j];
    r2.l[j] = FR_R2(j);
  }
  bool good = limbs_gt(m, l, 8);
  if (good)
    fr::mul(r, raw, r2);
  else
    r.set_zero();
  out[i] = r;
  ok[i] = good ? 1 : 0;
}
// Fp (the 12 coefficients of a GT value, tower order): 48 B LE each
__global__ void k_fp_to_bytes(const fp* __restrict__ in, uint32_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t c[12];
  fp_canon(c, in[i]);
  for (int j = 0; j < 12; j++) out[i * 12 + j] = c[j];
}
__global__ void k_fp_from_bytes(const uint32_t* __restrict__ in, fp* __restrict__ out, uint8_t* __restrict__ ok, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t l[12], m[12];
  fp raw, r2, r;
  for (int j = 0; j < 12; j++) {
    l[j] = in[i * 12 + j];
    m[j] = FpParams::mod(j);
    raw.l[j] = l[j];
    r2.l[j] = FP_R2(j);
  }
  bool good = limbs_gt(m, l, 12);
  if (good)
    fp::mul(r, raw, r2);
  else
    r.set_zero();
  out[i] = r;
  ok[i] = good ? 1 : 0;
}
// PairingOutput's Valid::check (ark-ec 0.5, what deserialize with Validate::Yes runs on a GT value): f^r == 1, i.e. f is
// in the order-r subgroup; zero and every other Fp12 value are rejected.  thread -> one GT value (tower code: 254
// squarings + the products of r's set bits; a CRS or an equation carries one GT value, so this is not a hot path).
__global__ void __launch_bounds__(64) k_gt_check_order(const fp12* __restrict__ in, const uint8_t* __restrict__ ok12,
                                                       fp12* __restrict__ out, uint8_t* __restrict__ ok, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool good = true;
  for (int j = 0; j < 12; j++) good = good && ok12[i * 12 + j];
  if (good) {
    const fp12 f = in[i];
    fp12 acc = f;
    for (int bit = 253; bit >= 0; bit--) {          // r has 255 bits: bit 254 is the leading one
      fp12::sqr(acc, acc);
      if ((FrParams::mod(bit >> 5) >> (bit & 31)) & 1) fp12::mul(acc, acc, f);
    }
    fp12 one;
    one.set_one();
    good = acc.equals(one);
  }
  if (!good) {
    fp* o = (fp*)&out[i];
    for (int j = 0; j < 12; j++) o[j].set_zero();
  }
  ok[i] = good ? 1 : 0;
}

}  // namespace gs

namespace {

// in: n elements of `isz` bytes on the host -> kernel -> n elements of `osz` bytes (+ n verdict bytes) back
template <class Launch>
int convert(gs_ctx* ctx, size_t n, const void* in, size_t isz, void* out, size_t osz, uint8_t* out_ok, Launch&& launch) {
  if (!ctx || (n && (!in || !out))) return GS_EARG;
  if (n == 0) return GS_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  uint8_t *din, *dout, *dok = nullptr;
  CUDA_TRY(upload(ctx, sc, &din, in, n * isz));
  CUDA_TRY(sc.alloc(&dout, n * osz));
  if (out_ok) CUDA_TRY(sc.alloc(&dok, n));
  int rc = launch(din, dout, dok);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(out, dout, n * osz, cudaMemcpyDeviceToHost, ctx->stream));
  if (out_ok) CUDA_TRY(cudaMemcpyAsync(out_ok, dok, n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

}  // namespace

extern "C" {

int gs_g1_compress(gs_ctx* ctx, size_t n, const gs_g1* pts, uint8_t* out) {
  return convert(ctx, n, pts, sizeof(gs_g1), out, 48, nullptr, [&](uint8_t* di, uint8_t* dout, uint8_t*) {
    LAUNCH(k_g1_compress, n, (const g1_aff*)di, dout, n);
    return GS_OK;
  });
}
int gs_g1_decompress(gs_ctx* ctx, size_t n, const uint8_t* in, int check_subgroup, gs_g1* out, uint8_t* out_ok) {
  if (n && !out_ok) return GS_EARG;
  return convert(ctx, n, in, 48, out, sizeof(gs_g1), out_ok, [&](uint8_t* di, uint8_t* dout, uint8_t* dok) {
    LAUNCH(k_g1_decompress, n, di, (g1_aff*)dout, dok, n, check_subgroup);
    return GS_OK;
  });
}
int gs_g2_compress(gs_ctx* ctx, size_t n, const gs_g2* pts, uint8_t* out) {
  return convert(ctx, n, pts, sizeof(gs_g2), out, 96, nullptr, [&](uint8_t* di, uint8_t* dout, uint8_t*) {
    LAUNCH(k_g2_compress, n, (const g2_aff*)di, dout, n);
    return GS_OK;
  });
}
int gs_g2_decompress(gs_ctx* ctx, size_t n, const uint8_t* in, int check_subgroup, gs_g2* out, uint8_t* out_ok) {
  if (n && !out_ok) return GS_EARG;
  return convert(ctx, n, in, 96, out, sizeof(gs_g2), out_ok, [&](uint8_t* di, uint8_t* dout, uint8_t* dok) {
    LAUNCH(k_g2_decompress, n, di, (g2_aff*)dout, dok, n, check_subgroup);
    return GS_OK;
  });
}
int gs_g1_serialize_uncompressed(gs_ctx* ctx, size_t n, const gs_g1* pts, uint8_t* out) {
  return convert(ctx, n, pts, sizeof(gs_g1), out, 96, nullptr, [&](uint8_t* di, uint8_t* dout, uint8_t*) {
    LAUNCH(k_g1_uncompressed, n, (const g1_aff*)di, dout, n);
    return GS_OK;
  });
}
int gs_g1_deserialize_uncompressed(gs_ctx* ctx, size_t n, const uint8_t* in, int check_subgroup, gs_g1* out, uint8_t* out_ok) {
  if (n && !out_ok) return GS_EARG;
  return convert(ctx, n, in, 96, out, sizeof(gs_g1), out_ok, [&](uint8_t* di, uint8_t* dout, uint8_t* dok) {
    LAUNCH(k_g1_from_uncompressed, n, di, (g1_aff*)dout, dok, n, check_subgroup);
    return GS_OK;
  });
}
int gs_g2_serialize_uncompressed(gs_ctx* ctx, size_t n, const gs_g2* pts, uint8_t* out) {
  return convert(ctx, n, pts, sizeof(gs_g2), out, 192, nullptr, [&](uint8_t* di, uint8_t* dout, uint8_t*) {
    LAUNCH(k_g2_uncompressed, n, (const g2_aff*)di, dout, n);
    return GS_OK;
  });
}
int gs_g2_deserialize_uncompressed(gs_ctx* ctx, size_t n, const uint8_t* in, int check_subgroup, gs_g2* out, uint8_t* out_ok) {
  if (n && !out_ok) return GS_EARG;
  return convert(ctx, n, in, 192, out, sizeof(gs_g2), out_ok, [&](uint8_t* di, uint8_t* dout, uint8_t* dok) {
    LAUNCH(k_g2_from_uncompressed, n, di, (g2_aff*)dout, dok, n, check_subgroup);
    return GS_OK;
  });
}
int gs_fr_to_bytes(gs_ctx* ctx, size_t n, const gs_fr* in, uint8_t* out) {
  return convert(ctx, n, in, 32, out, 32, nullptr, [&](uint8_t* di, uint8_t* dout, uint8_t*) {
    LAUNCH(k_fr_to_bytes, n, (const fr*)di, (uint32_t*)dout, n);
    return GS_OK;
  });
}
int gs_fr_from_bytes(gs_ctx* ctx, size_t n, const uint8_t* in, gs_fr* out, uint8_t* out_ok) {
  if (n && !out_ok) return GS_EARG;
  return convert(ctx, n, in, 32, out, 32, out_ok, [&](uint8_t* di, uint8_t* dout, uint8_t* dok) {
    LAUNCH(k_fr_from_bytes, n, (const uint32_t*)di, (fr*)dout, dok, n);
    return GS_OK;
  });
}
int gs_gt_to_bytes(gs_ctx* ctx, size_t n, const gs_gt* in, uint8_t* out) {
  return convert(ctx, n, in, 576, out, 576, nullptr, [&](uint8_t* di, uint8_t* dout, uint8_t*) {
    LAUNCH(k_fp_to_bytes, n * 12, (const fp*)di, (uint32_t*)dout, n * 12);
    return GS_OK;
  });
}
int gs_gt_from_bytes(gs_ctx* ctx, size_t n, const uint8_t* in, gs_gt* out, uint8_t* out_ok) {
  if (n && !out_ok) return GS_EARG;
  // every coefficient canonical (< p), then PairingOutput's Valid::check: f^r == 1 (k_gt_check_order)
  if (!ctx || (n && (!in || !out))) return GS_EARG;
  if (n == 0) return GS_OK;
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  uint8_t *din, *dok12, *dok;
  fp* dout;
  CUDA_TRY(upload(ctx, sc, &din, in, n * 576));
  CUDA_TRY(sc.alloc(&dout, n * 12));
  CUDA_TRY(sc.alloc(&dok12, n * 12));
  CUDA_TRY(sc.alloc(&dok, n));
  LAUNCH(k_fp_from_bytes, n * 12, (const uint32_t*)din, dout, dok12, n * 12);
  LAUNCH_CFG(k_gt_check_order, n, 64, 0, (const fp12*)dout, dok12, (fp12*)dout, dok, n);
  CUDA_TRY(cudaMemcpyAsync(out, dout, n * 576, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(out_ok, dok, n, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return GS_OK;
}

}  // extern "C"
