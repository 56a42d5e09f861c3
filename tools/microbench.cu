// Register-only microbenchmarks that calibrate the integer-multiply roofline on the box:
//   (1) raw IMAD.WIDE.U32 issue rate (independent chains)
//   (2) Fp Montgomery products per second (fp::mul chains), for several occupancies
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../groth-sahai-rs_b200/csrc/fp.cuh"
#include "../groth-sahai-rs_b200/csrc/tower.cuh"
#include "../groth-sahai-rs_b200/csrc/modinv.cuh"
#include "experimental/fpd.cuh"
using namespace gs;

__global__ void k_imad_wide(unsigned long long* out, unsigned a, unsigned b, int iters) {
  unsigned long long x0 = threadIdx.x, x1 = a, x2 = b, x3 = a + b, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
  unsigned m = a | 1;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x0) : "r"(m), "r"((unsigned)x1));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x1) : "r"(m), "r"((unsigned)x2));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x2) : "r"(m), "r"((unsigned)x3));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x3) : "r"(m), "r"((unsigned)x4));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x4) : "r"(m), "r"((unsigned)x5));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x5) : "r"(m), "r"((unsigned)x6));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x6) : "r"(m), "r"((unsigned)x7));
      asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x7) : "r"(m), "r"((unsigned)x0));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}
__global__ void k_imad_lo(unsigned* out, unsigned a, unsigned b, int iters) {
  unsigned x0 = threadIdx.x, x1 = a, x2 = b, x3 = a + b, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
  unsigned m = a | 1;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x0) : "r"(m), "r"(x1));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x1) : "r"(m), "r"(x2));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x2) : "r"(m), "r"(x3));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x3) : "r"(m), "r"(x4));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x4) : "r"(m), "r"(x5));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x5) : "r"(m), "r"(x6));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x6) : "r"(m), "r"(x7));
      asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x7) : "r"(m), "r"(x0));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}
template <int ILP>
__global__ void k_fpmul(fp* out, const fp* in, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  fp x[ILP], y = in[t];
  for (int j = 0; j < ILP; j++) x[j] = in[t + j + 1];
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) fp::mul(x[j], x[j], y);
  }
  fp r = x[0];
  for (int j = 1; j < ILP; j++) fp::add(r, r, x[j]);
  out[t] = r;
}
__global__ void k_fpinv_fermat(fp* out, const fp* in, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  fp x = in[t];
  for (int i = 0; i < iters; i++) fp_inv(x, x);
  out[t] = x;
}
__global__ void k_fpinv_safegcd(fp* out, const fp* in, int iters) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  fp x = in[t];
  for (int i = 0; i < iters; i++) fp_inv_sg(x, x);
  out[t] = x;
}
// safegcd inversions on half of the warps next to fp::mul chains on the other half: do the pipes overlap?
__global__ void k_mix(fp* out, const fp* in, int iters_inv, int iters_mul) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  fp x = in[t], y = in[t + 1];
  if ((threadIdx.x >> 5) & 1) {
    for (int i = 0; i < iters_inv; i++) fp_inv_sg(x, x);
  } else {
    for (int i = 0; i < iters_mul; i++) fp::mul(x, x, y);
  }
  out[t] = x;
}
// FP64 pipe: independent DFMA chains; and DFMA on half of the warps next to IMAD.WIDE on the other half
__global__ void k_dfma(double* out, double a, double b, int iters) {
  double x0 = threadIdx.x, x1 = a, x2 = b, x3 = a + b, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < 16; j++) {
      x0 = __fma_rz(x0, a, b); x1 = __fma_rz(x1, a, b); x2 = __fma_rz(x2, a, b); x3 = __fma_rz(x3, a, b);
      x4 = __fma_rz(x4, a, b); x5 = __fma_rz(x5, a, b); x6 = __fma_rz(x6, a, b); x7 = __fma_rz(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void k_dfma_imad_mix(unsigned long long* out, double a, double b, unsigned ua, int it_d, int it_i) {
  if ((threadIdx.x >> 5) & 1) {
    double x0 = threadIdx.x, x1 = a, x2 = b, x3 = a + b, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
    for (int i = 0; i < it_d; i++) {
#pragma unroll
      for (int j = 0; j < 16; j++) {
        x0 = __fma_rz(x0, a, b); x1 = __fma_rz(x1, a, b); x2 = __fma_rz(x2, a, b); x3 = __fma_rz(x3, a, b);
        x4 = __fma_rz(x4, a, b); x5 = __fma_rz(x5, a, b); x6 = __fma_rz(x6, a, b); x7 = __fma_rz(x7, a, b);
      }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = (unsigned long long)(x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7);
  } else {
    unsigned long long x0 = threadIdx.x, x1 = ua, x2 = 3, x3 = 4, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
    unsigned m = ua | 1;
    for (int i = 0; i < it_i; i++) {
#pragma unroll
      for (int j = 0; j < 16; j++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x0) : "r"(m), "r"((unsigned)x1));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x1) : "r"(m), "r"((unsigned)x2));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x2) : "r"(m), "r"((unsigned)x3));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x3) : "r"(m), "r"((unsigned)x4));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x4) : "r"(m), "r"((unsigned)x5));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x5) : "r"(m), "r"((unsigned)x6));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x6) : "r"(m), "r"((unsigned)x7));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x7) : "r"(m), "r"((unsigned)x0));
      }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
  }
}
// sums of 4 Fp products with one Montgomery reduction: integer pipe (fp::mulsum) vs FP64 pipe (fpd.cuh), and
// both at once on alternating warps.  mode: 0 = all warps integer, 1 = all warps FP64, 2 = odd warps FP64;
// `it_i` / `it_d` = chain length of the integer / FP64 warps (0 switches that half off).
template <int mode>
__global__ void __launch_bounds__(128) k_mulsum_mix(fp* out, const fp* in, int it_i, int it_d) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  fp a[4], b[4];
  for (int j = 0; j < 4; j++) {
    a[j] = in[t + j];
    b[j] = in[t + 4 + j];
  }
  const bool dfma = mode == 1 || (mode == 2 && ((threadIdx.x >> 5) & 1));
  fp r = a[0];
  if (dfma) {
    for (int i = 0; i < it_d; i++) {
      mulsum_dfma<4>(r, a, b);
      a[i & 3] = r;
    }
  } else {
    for (int i = 0; i < it_i; i++) {
      fp::mulsum<4>(r, a, b);
      a[i & 3] = r;
    }
  }
  out[t] = r;
}
// the same with the ROLLED FP64 form (b operands in shared memory, word-major / thread-minor):
// mode 3 = all warps FP64 rolled, 4 = odd warps FP64 rolled + even warps integer
template <int mode>
__global__ void __launch_bounds__(128) k_mulsum_mix_r(fp* out, const fp* in, int it_i, int it_d) {
  __shared__ uint32_t sb[4 * 12 * 128];
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  fp a[4], b[4];
  for (int j = 0; j < 4; j++) {
    a[j] = in[t + j];
    b[j] = in[t + 4 + j];
    for (int w = 0; w < 12; w++) sb[(j * 12 + w) * 128 + threadIdx.x] = b[j].l[w];
  }
  __syncthreads();
  const bool dfma = mode == 3 || (mode == 4 && ((threadIdx.x >> 5) & 1));
  fp r = a[0];
  if (dfma) {
    for (int i = 0; i < it_d; i++) {
      mulsum_dfma_rolled<4>(r, a, [&](int tt, int w) { return (uint64_t)sb[(tt * 12 + w) * 128 + threadIdx.x]; });
      a[i & 3] = r;
    }
  } else {
    for (int i = 0; i < it_i; i++) {
      fp::mulsum<4>(r, a, b);
      a[i & 3] = r;
    }
  }
  out[t] = r;
}
template <class K, class... A>
float timeit(K k, dim3 g, dim3 b, A... args) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<g, b>>>(args...); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<<<g, b>>>(args...); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  void* buf; cudaMalloc(&buf, 1 << 28); cudaMemset(buf, 1, 1 << 28);
  for (int wps : {4, 8, 16, 32}) {  // warps per SM
    int iters = 4096;
    dim3 g(sms), b(wps * 32);
    float ms = timeit(k_imad_wide, g, b, (unsigned long long*)buf, 3u, 5u, iters);
    double ops = (double)sms * wps * 32 * iters * 128;
    float ms2 = timeit(k_imad_lo, g, b, (unsigned*)buf, 3u, 5u, iters);
    printf("warps/SM %2d: IMAD.WIDE %.3e/s (%.1f /clk/SM @1.9GHz)   IMAD.lo %.3e/s (%.1f /clk/SM)\n", wps, ops / ms * 1e3,
           ops / ms * 1e3 / sms / 1.9e9, ops / ms2 * 1e3, ops / ms2 * 1e3 / sms / 1.9e9);
  }
  for (int wps : {4, 8, 12, 16, 24, 32}) {
    int iters = 2000;
    dim3 g(sms), b(wps * 32);
    float m1 = timeit(k_fpmul<1>, g, b, (fp*)buf + (1 << 20), (const fp*)buf, iters);
    float m2 = timeit(k_fpmul<2>, g, b, (fp*)buf + (1 << 20), (const fp*)buf, iters);
    double n1 = (double)sms * wps * 32 * iters, n2 = n1 * 2;
    printf("warps/SM %2d: fp::mul ILP1 %.3e M/s   ILP2 %.3e M/s\n", wps, n1 / m1 * 1e3, n2 / m2 * 1e3);
  }
  for (int wps : {4, 8, 16, 32}) {
    dim3 g(sms), b(wps * 32);
    int iters = 8;
    float m1 = timeit(k_fpinv_fermat, g, b, (fp*)buf + (1 << 20), (const fp*)buf, iters);
    float m2 = timeit(k_fpinv_safegcd, g, b, (fp*)buf + (1 << 20), (const fp*)buf, iters);
    double n = (double)sms * wps * 32 * iters;
    printf("warps/SM %2d: fp_inv Fermat %.3e inv/s (%.0f M-equivalents each)   safegcd %.3e inv/s (%.0f M-equivalents each)\n", wps,
           n / m1 * 1e3, 3.08e10 / (n / m1 * 1e3), n / m2 * 1e3, 3.08e10 / (n / m2 * 1e3));
  }
  for (int wps : {8, 16, 32}) {
    dim3 g(sms), b(wps * 32);
    int ii = 8, im = 8 * 120;
    float ma = timeit(k_mix, g, b, (fp*)buf + (1 << 20), (const fp*)buf, ii, 0);
    float mb = timeit(k_mix, g, b, (fp*)buf + (1 << 20), (const fp*)buf, 0, im);
    float mc = timeit(k_mix, g, b, (fp*)buf + (1 << 20), (const fp*)buf, ii, im);
    printf("warps/SM %2d: half warps safegcd alone %.3f ms, half warps fp::mul alone %.3f ms, both %.3f ms\n", wps, ma, mb, mc);
  }
  for (int wps : {4, 8, 12, 16}) {  // 4-warp blocks, wps / 4 blocks per SM (as many as the registers allow)
    dim3 g(sms * (wps / 4)), b(128);
    int it = 400;
    fp* o = (fp*)buf + (1 << 20);
    float mi = timeit(k_mulsum_mix<0>, g, b, o, (const fp*)buf, it, it);
    float md = timeit(k_mulsum_mix<1>, g, b, o, (const fp*)buf, it, it);
    float hi = timeit(k_mulsum_mix<2>, g, b, o, (const fp*)buf, it, 0);
    float hd = timeit(k_mulsum_mix<2>, g, b, o, (const fp*)buf, 0, it);
    float hb = timeit(k_mulsum_mix<2>, g, b, o, (const fp*)buf, it, it);
    double n = (double)sms * wps * 32 * it;
    printf("warps/SM %2d: mulsum<4> integer %.3e/s   FP64 %.3e/s   | half warps: integer alone %.3f ms, FP64 alone %.3f ms, both %.3f ms "
           "=> %.3e/s combined\n", wps, n / mi * 1e3, n / md * 1e3, hi, hd, hb, n / hb * 1e3);
    float rd = timeit(k_mulsum_mix_r<3>, g, b, o, (const fp*)buf, it, it);
    float ri = timeit(k_mulsum_mix_r<4>, g, b, o, (const fp*)buf, it, 0);
    float rdd = timeit(k_mulsum_mix_r<4>, g, b, o, (const fp*)buf, 0, it);
    float rb = timeit(k_mulsum_mix_r<4>, g, b, o, (const fp*)buf, it, it);
    printf("             rolled FP64 form: all warps %.3e/s   | half warps: integer alone %.3f ms, FP64 alone %.3f ms, both %.3f ms "
           "=> %.3e/s combined\n", n / rd * 1e3, ri, rdd, rb, n / rb * 1e3);
  }
  for (int wps : {8, 16, 32}) {
    int iters = 4096;
    dim3 g(sms), b(wps * 32);
    float ms = timeit(k_dfma, g, b, (double*)buf, 1.0000001, 0.5, iters);
    double ops = (double)sms * wps * 32 * iters * 128;
    printf("warps/SM %2d: DFMA %.3e/s (%.1f /clk/SM @1.9GHz)\n", wps, ops / ms * 1e3, ops / ms * 1e3 / sms / 1.9e9);
    float ma = timeit(k_dfma_imad_mix, g, b, (unsigned long long*)buf, 1.0000001, 0.5, 3u, iters, 0);
    float mb = timeit(k_dfma_imad_mix, g, b, (unsigned long long*)buf, 1.0000001, 0.5, 3u, 0, iters);
    float mc = timeit(k_dfma_imad_mix, g, b, (unsigned long long*)buf, 1.0000001, 0.5, 3u, iters, iters);
    printf("warps/SM %2d: half warps DFMA alone %.3f ms, half warps IMAD.WIDE alone %.3f ms, both %.3f ms\n", wps, ma, mb, mc);
  }
  return 0;
}
