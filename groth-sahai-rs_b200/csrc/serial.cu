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
#include "endo.cuh"

using namespace gs;

namespace gs {

static __device__ __constant__ uint32_t EXP_P14[12] = {0xffffeaabu, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                                  0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};   // (p+1)/4
static __device__ __constant__ uint32_t EXP_PM34[12] = {0xffffeaaau, 0xee7fbfffu, 0xac54ffffu, 0x07aaffffu, 0x3dac3d89u, 0xd9cc34a8u,
                                   0x3ce144afu, 0xd91dd2e1u, 0x90d2eb35u, 0x92c6e9edu, 0x8e5ff9a6u, 0x0680447au};  // (p-3)/4
static __device__ __constant__ uint32_t EXP_PM12[12] = {0xffffd555u, 0xdcff7fffu, 0x58a9ffffu, 0x0f55ffffu, 0x7b587b12u, 0xb3986950u,
                                   0x79c2895fu, 0xb23ba5c2u, 0x21a5d66bu, 0x258dd3dbu, 0x1cbff34du, 0x0d0088f5u};  // (p-1)/2

enum { EXP_SEL_P14 = 0, EXP_SEL_PM34 = 1, EXP_SEL_PM12 = 2 };
__device__ GS_INL uint32_t exp_limb(int sel, int i) {
  return sel == EXP_SEL_P14 ? EXP_P14[i] : (sel == EXP_SEL_PM34 ? EXP_PM34[i] : EXP_PM12[i]);
}

// r = a^e, e one of the three 381-bit constants above (left-to-right binary; the exponents are public)
template <class F>
__device__ GS_NOINL void pow_const(typename F::T& r, const typename F::T& a, int sel) {
  typename F::T acc;
  F::set_one(acc);
  bool started = false;
#pragma unroll 1
  for (int i = 11; i >= 0; i--) {
    const uint32_t w = exp_limb(sel, i);
#pragma unroll 1
    for (int b = 31; b >= 0; b--) {
      if (started) F::sqr(acc, acc);
      if ((w >> b) & 1) {
        if (started)
          F::mul(acc, acc, a);
        else
          acc = a;
        started = true;
      }
    }
  }
  r = acc;
}

// canonical (non-Montgomery) limbs of a
__device__ GS_INL void fp_canon(uint32_t out[12], const fp& a) {
  fp one_raw, t;
  one_raw.set_zero();
  one_raw.l[0] = 1;
  fp::mul(t, a, one_raw);
#pragma unroll
  for (int i = 0; i < 12; i++) out[i] = t.l[i];
}
__device__ GS_INL bool limbs_gt(const uint32_t* a, const uint32_t* b, int n) {  // a > b
  for (int i = n - 1; i >= 0; i--) {
    if (a[i] != b[i]) return a[i] > b[i];
  }
  return false;
}
__device__ GS_INL bool fp_is_largest(const fp& y) {  // y > (p-1)/2  <=>  y > -y as integers
  uint32_t c[12];
  fp_canon(c, y);
  return limbs_gt(c, EXP_PM12, 12);
}
__device__ GS_INL bool fp2_is_largest(const fp2& y) {  // Fp2 ordered with c1 most significant (ark-ff Ord, zcash)
  if (!y.c1.is_zero()) return fp_is_largest(y.c1);
  return fp_is_largest(y.c0);
}
// 48 big-endian bytes (top three bits of byte 0 masked off) -> Montgomery Fp; false when the integer is >= p
__device__ GS_INL bool fp_from_be(fp& r, const uint8_t* b) {
  uint32_t l[12], m[12];
#pragma unroll
  for (int j = 0; j < 12; j++) {
    const uint8_t* q = b + 44 - 4 * j;
    uint32_t hi = q[0];
    if (j == 11) hi &= 0x1Fu;
    l[j] = (hi << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    m[j] = FpParams::mod(j);
  }
  if (!limbs_gt(m, l, 12)) return false;
  fp raw, r2;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    raw.l[j] = l[j];
    r2.l[j] = FP_R2(j);
  }
  fp::mul(r, raw, r2);
  return true;
}
__device__ GS_INL void fp_to_be(uint8_t* b, const fp& a) {
  uint32_t c[12];
  fp_canon(c, a);
#pragma unroll
  for (int j = 0; j < 12; j++) {
    uint8_t* q = b + 44 - 4 * j;
    q[0] = (uint8_t)(c[j] >> 24);
    q[1] = (uint8_t)(c[j] >> 16);
    q[2] = (uint8_t)(c[j] >> 8);
    q[3] = (uint8_t)c[j];
  }
}

__device__ GS_INL bool fp_sqrt(fp& y, const fp& a) {  // p = 3 mod 4
  pow_const<FpOps>(y, a, EXP_SEL_P14);
  fp t;
  fp::sqr(t, y);
  return t.equals(a);
}
// Adj - Rodriguez-Henriquez, "Square root computation over even extension fields", Alg. 9 (q = 3 mod 4)
__device__ GS_NOINL bool fp2_sqrt(fp2& x, const fp2& a) {
  fp2 a1, alpha, x0, t, minus_one;
  pow_const<Fp2Ops>(a1, a, EXP_SEL_PM34);
  fp2::mul(x0, a1, a);        // a^((q+1)/4)
  fp2::mul(alpha, a1, x0);    // a^((q-1)/2)
  minus_one.set_zero();
  fp_one(minus_one.c0);
  fp::neg(minus_one.c0, minus_one.c0);
  if (alpha.equals(minus_one)) {  // x = u * x0
    fp::neg(x.c0, x0.c1);
    x.c1 = x0.c0;
  } else {
    fp2 b;
    t = alpha;
    fp one;
    fp_one(one);
    fp::add(t.c0, t.c0, one);
    pow_const<Fp2Ops>(b, t, EXP_SEL_PM12);
    fp2::mul(x, b, x0);
  }
  fp2::sqr(t, x);
  return t.equals(a);  // also rejects non-residues (Alg. 9's a0 = -1 test)
}

// ------------------------------------------------------------------ subgroup membership (Scott, ePrint 2021/1130)
// The tests ark-bls12-381 runs in is_in_correct_subgroup_assuming_on_curve, 64-bit scalars instead of [r]P:
//   G1 (Section 6):  phi(P) = -[x^2] P,  phi(x, y) = (beta x, y);  additionally [x]P = P (P != O) is rejected
//   G2 (Section 4):  psi(Q) = [x] Q,     psi(x, y) = (conj(x) cx, conj(y) cy)  (untwist-Frobenius-twist)
// beta, cx, cy below were derived from those relations on the generators (tests/test_serialize.py checks the
// oracle versions against the definition [r]P = O on points inside and outside the subgroups).
// r = [|x|] b, |x| = 0xd201000000010000 (63 doublings, 5 additions)
template <class F>
__device__ GS_NOINL void mul_x_abs(Jac<F>& r, const Jac<F>& b) {
  Jac<F> acc = b;
#pragma unroll 1
  for (int bit = 62; bit >= 0; bit--) {
    Jac<F>::dbl(acc, acc);
    if ((0xd201000000010000ull >> bit) & 1) Jac<F>::add(acc, acc, b);
  }
  r = acc;
}
// Jacobian j == affine (ax, ay) ?   (j finite)
template <class F>
__device__ GS_INL bool jac_equals_affine(const Jac<F>& j, const typename F::T& ax, const typename F::T& ay) {
  if (j.is_inf()) return false;
  typename F::T z2, z3, t;
  F::sqr(z2, j.Z);
  F::mul(z3, z2, j.Z);
  F::mul(t, ax, z2);
  if (!t.equals(j.X)) return false;
  F::mul(t, ay, z3);
  return t.equals(j.Y);
}
__device__ GS_NOINL bool in_subgroup_g1(const g1_aff& p) {
  g1_jac b, t1, t2;
  b.from_affine(p);
  mul_x_abs<FpOps>(t1, b);
  if (jac_equals_affine<FpOps>(t1, p.x, p.y)) return false;  // [x]P = P
  mul_x_abs<FpOps>(t2, t1);                                  // [x^2] P
  fp bx, ny;
  endo_phi_x(bx, p.x);
  fp::neg(ny, p.y);
  return jac_equals_affine<FpOps>(t2, bx, ny);               // [x^2]P = -phi(P)
}
__device__ GS_NOINL bool in_subgroup_g2(const g2_aff& q) {
  g2_jac b, t;
  b.from_affine(q);
  mul_x_abs<Fp2Ops>(t, b);                                   // [|x|] Q = -[x] Q
  g2_aff ps;
  endo_psi(ps, q);
  fp2 px = ps.x, py;
  fp2::neg(py, ps.y);
  return jac_equals_affine<Fp2Ops>(t, px, py);               // [|x|]Q = -psi(Q)
}

// ------------------------------------------------------------------ G1
__global__ void __launch_bounds__(128) k_g1_compress(const g1_aff* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  g1_aff p = in[i];
  uint8_t b[48];
  if (p.is_inf()) {
    for (int j = 0; j < 48; j++) b[j] = 0;
    b[0] = 0xC0;
  } else {
    fp_to_be(b, p.x);
    b[0] |= 0x80 | (fp_is_largest(p.y) ? 0x20 : 0);
  }
  for (int j = 0; j < 48; j++) out[i * 48 + j] = b[j];
}
__global__ void __launch_bounds__(128) k_g1_decompress(const uint8_t* __restrict__ in, g1_aff* __restrict__ out, uint8_t* __restrict__ ok,
                                                       size_t n, int check_subgroup) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[48];
  for (int j = 0; j < 48; j++) b[j] = in[i * 48 + j];
  g1_aff p;
  p.set_inf();
  bool good = (b[0] & 0x80) != 0;
  if (good && !(b[0] & 0x40)) {
    good = fp_from_be(p.x, b);
    if (good) {
      fp rhs, four;
      fp::sqr(rhs, p.x);
      fp::mul(rhs, rhs, p.x);
      for (int j = 0; j < 12; j++) four.l[j] = FP_FOUR(j);
      fp::add(rhs, rhs, four);
      good = fp_sqrt(p.y, rhs);
      if (good) {
        if (fp_is_largest(p.y) != ((b[0] & 0x20) != 0)) fp::neg(p.y, p.y);
        if (check_subgroup) good = in_subgroup_g1(p);
      }
    }
    if (!good) p.set_inf();
  }
  out[i] = p;
  ok[i] = good ? 1 : 0;
}

// ------------------------------------------------------------------ G2
__global__ void __launch_bounds__(128) k_g2_compress(const g2_aff* __restrict__ in, uint8_t* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  g2_aff p = in[i];
  uint8_t b[96];
  if (p.is_inf()) {
    for (int j = 0; j < 96; j++) b[j] = 0;
    b[0] = 0xC0;
  } else {
    fp_to_be(b, p.x.c1);
    fp_to_be(b + 48, p.x.c0);
    b[0] |= 0x80 | (fp2_is_largest(p.y) ? 0x20 : 0);
  }
  for (int j = 0; j < 96; j++) out[i * 96 + j] = b[j];
}
__global__ void __launch_bounds__(128) k_g2_decompress(const uint8_t* __restrict__ in, g2_aff* __restrict__ out, uint8_t* __restrict__ ok,
                                                       size_t n, int check_subgroup) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t b[96];
  for (int j = 0; j < 96; j++) b[j] = in[i * 96 + j];
  g2_aff p;
  p.set_inf();
  bool good = (b[0] & 0x80) != 0;
  if (good && !(b[0] & 0x40)) {
    good = fp_from_be(p.x.c1, b);
    uint8_t save = b[48];
    if (good) {
      // the second coordinate has no flag bits: a set top bit means >= p
      good = (save & 0xE0) == 0 && fp_from_be(p.x.c0, b + 48);
    }
    if (good) {
      fp2 rhs, bt;
      fp2::sqr(rhs, p.x);
      fp2::mul(rhs, rhs, p.x);
      for (int j = 0; j < 12; j++) bt.c0.l[j] = bt.c1.l[j] = FP_FOUR(j);  // b' = 4 (1 + u)
      fp2::add(rhs, rhs, bt);
      good = fp2_sqrt(p.y, rhs);
      if (good) {
        if (fp2_is_largest(p.y) != ((b[0] & 0x20) != 0)) fp2::neg(p.y, p.y);
        if (check_subgroup) good = in_subgroup_g2(p);
      }
    }
    if (!good) p.set_inf();
  }
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
  // one verdict byte per Fp coefficient on the device, folded to one per GT value on the host
  if (!ctx || (n && (!in || !out))) return GS_EARG;
  if (n == 0) return GS_OK;
  std::vector<uint8_t> ok12(n * 12);
  CUDA_TRY(cudaSetDevice(ctx->device));
  Scratch sc(ctx);
  uint8_t *din, *dok;
  fp* dout;
  CUDA_TRY(upload(ctx, sc, &din, in, n * 576));
  CUDA_TRY(sc.alloc(&dout, n * 12));
  CUDA_TRY(sc.alloc(&dok, n * 12));
  LAUNCH(k_fp_from_bytes, n * 12, (const uint32_t*)din, dout, dok, n * 12);
  CUDA_TRY(cudaMemcpyAsync(out, dout, n * 576, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaMemcpyAsync(ok12.data(), dok, n * 12, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (size_t i = 0; i < n; i++) {
    uint8_t g = 1;
    for (int j = 0; j < 12; j++) g &= ok12[i * 12 + j];
    out_ok[i] = g;
  }
  return GS_OK;
}

}  // extern "C"
