// Bucket (Pippenger) multi-scalar multiplication for the proof elements of ONE large statement
// (north star: "Pippenger-style G1/G2 MSMs for the pi / theta proof elements"; replaces the term-by-term
// `left_mul(.., true)` of src/prover/prove.rs:129-160 / src/data_structures.rs:696-742 when the MSM is big).
//
//   out[row] = sum_t sv[row][t] * B_t          rows <= 2 scalar rows over the same N bases (pi_1, pi_2 / theta_1, theta_2)
//
// 1. k_pip_expand   every term is split along the group's endomorphism (endo.cuh: 2 x 128-bit sub-scalars on G1,
//                   4 x 64-bit on G2) so that the windows only span 128 / 64 bits: Np = N * PARTS points
//                   (-1)^j endo^j(B_t), and signed c-bit digits d in [-H, H], H = 2^(c-1), for W windows each (pip_choose).
// 2. k_pip_hist / k_pip_scan / k_pip_scatter   counting sort of the points of every (row, window) by |digit|.
// 3. k_pip_accumulate   thread = (row, window, bucket): sum of its segment, mixed additions (the only O(Np) step:
//                   W * Np additions per row instead of ~1.3 * 255 per term in the windowed scalar multiplications).
// 4. k_pip_bucket_reduce   S_w = sum_b (b+1) Bucket_b by running sums over groups of G buckets, a small scalar
//                   multiplication for the group offset, then the pairwise tree of reduce_rows.
// 5. k_pip_window_shift + tree   out = sum_w 2^(c w) S_w.
// Every step is deterministic up to the ORDER of additions inside a bucket (atomics in the scatter); the result is a
// group element, normalised to affine by k_proof_finish, so the bytes do not depend on it.
#pragma once
// (included by prover_impl.cuh after reduce_rows)
#include "ctx.h"
#include "endo.cuh"

namespace gs {

struct pip_geom {
  int c, W, H;   // window bits, windows, buckets per window (2^(c-1))
  int G, NG;     // buckets per reduction group, groups per window
  int parts, bits;
  size_t N, Np;  // terms, points after the endomorphism split
};

// sub-scalars of k along the endomorphism as PARTS x uint32[4] (little-endian), each < 2^bits
template <class F>
struct PipSplit;
template <>
struct PipSplit<FpOps> {
  static constexpr int PARTS = 2, BITS = 128;
  GS_HD static GS_INL void split(uint32_t out[2][4], const uint32_t k[8]) { glv_split(out[0], out[1], k); }
  // image j of the base: j = 0: B, j = 1: -phi(B) = (beta x, -y)
  GS_HD static GS_INL void image(g1_aff& r, const g1_aff& b, int j) {
    r = b;
    if (j == 1 && !b.is_inf()) {
      endo_phi_x(r.x, b.x);
      fp::neg(r.y, b.y);
    }
  }
};
template <>
struct PipSplit<Fp2Ops> {
  static constexpr int PARTS = 4, BITS = 64;
  GS_HD static GS_INL void split(uint32_t out[4][4], const uint32_t k[8]) {
    uint64_t c[4];
    gls_split(c, k);
    for (int j = 0; j < 4; j++) {
      out[j][0] = (uint32_t)c[j];
      out[j][1] = (uint32_t)(c[j] >> 32);
      out[j][2] = out[j][3] = 0;
    }
  }
  // image j: (-1)^j psi^j(B)   (k = sum_j c_j |x|^j and psi = [x] = -[|x|])
  GS_HD static GS_INL void image(g2_aff& r, const g2_aff& b, int j) {
    r = b;
    for (int i = 0; i < j; i++) {
      g2_aff t;
      endo_psi(t, r);
      r = t;
    }
    if ((j & 1) && !r.is_inf()) fp2::neg(r.y, r.y);
  }
};

GS_HD GS_INL uint32_t pip_bits(const uint32_t k[4], int bit, int c) {
  int w = bit >> 5, s = bit & 31;
  if (w >= 4) return 0;
  uint64_t v = k[w];
  if (w + 1 < 4) v |= (uint64_t)k[w + 1] << 32;
  return (uint32_t)(v >> s) & ((1u << c) - 1u);
}

// thread -> term t: the PARTS images of its base and, for every row, the signed digits of the PARTS sub-scalars
// digits[((row*W + w) * Np) + t*PARTS + j]  (int16: |d| <= H <= 4096)
template <class F>
__global__ void __launch_bounds__(128) k_pip_expand(const Aff<F>* __restrict__ b0, size_t n0, const Aff<F>* __restrict__ b1, size_t n1,
                                                    const fr* __restrict__ sv, int rows, Aff<F>* __restrict__ pts,
                                                    int16_t* __restrict__ digits, pip_geom g) {
  constexpr int PARTS = PipSplit<F>::PARTS;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= g.N) return;
  const Aff<F> B = t < n0 ? b0[t] : b1[t - n0];
  for (int j = 0; j < PARTS; j++) {
    Aff<F> im;
    PipSplit<F>::image(im, B, j);
    pts[t * PARTS + j] = im;
  }
  const uint32_t half = 1u << (g.c - 1);
  for (int row = 0; row < rows; row++) {
    uint32_t k[8], sub[PARTS][4];
    fr_from_mont(k, sv[(size_t)row * g.N + t]);
    PipSplit<F>::split(sub, k);
    for (int j = 0; j < PARTS; j++) {
      uint32_t carry = 0;
      for (int w = 0; w < g.W; w++) {
        uint32_t d = pip_bits(sub[j], w * g.c, g.c) + carry;
        int sd;
        if (d > half) {
          sd = (int)d - (int)(1u << g.c);
          carry = 1;
        } else {
          sd = (int)d;
          carry = 0;
        }
        if (B.is_inf()) sd = 0;
        digits[((size_t)row * g.W + w) * g.Np + t * PARTS + j] = (int16_t)sd;
      }
    }
  }
}

// (the three sorting kernels do not depend on the group; they are templates only so that each lives in the TU of its group)
// thread -> (rw, q): counts[rw*H + |d|-1]++
template <class F>
__global__ void __launch_bounds__(128) k_pip_hist(const int16_t* __restrict__ digits, uint32_t* __restrict__ counts, size_t total, pip_geom g) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  int d = digits[id];
  if (d == 0) return;
  size_t rw = id / g.Np;
  atomicAdd(&counts[rw * g.H + (size_t)((d < 0 ? -d : d) - 1)], 1u);
}
// block -> rw: off[rw*(H+1) + b] = exclusive prefix sum of counts[rw*H ..]; cursor = copy of the offsets
template <class F>
__global__ void __launch_bounds__(256) k_pip_scan(const uint32_t* __restrict__ counts, uint32_t* __restrict__ off, uint32_t* __restrict__ cursor,
                                                  pip_geom g) {
  __shared__ uint32_t part[256];
  const size_t rw = blockIdx.x;
  const int per = (g.H + 255) / 256;
  const int b0 = threadIdx.x * per;
  uint32_t s = 0;
  for (int b = b0; b < b0 + per && b < g.H; b++) s += counts[rw * g.H + b];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < 256; i++) {
      uint32_t v = part[i];
      part[i] = run;
      run += v;
    }
    off[rw * (g.H + 1) + g.H] = run;
  }
  __syncthreads();
  uint32_t run = part[threadIdx.x];
  for (int b = b0; b < b0 + per && b < g.H; b++) {
    off[rw * (g.H + 1) + b] = run;
    cursor[rw * g.H + b] = run;
    run += counts[rw * g.H + b];
  }
}
// thread -> (rw, q): sorted[rw*Np + pos] = (q << 1) | negative
template <class F>
__global__ void __launch_bounds__(128) k_pip_scatter(const int16_t* __restrict__ digits, uint32_t* __restrict__ cursor,
                                                     uint32_t* __restrict__ sorted, size_t total, pip_geom g) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  int d = digits[id];
  if (d == 0) return;
  size_t rw = id / g.Np, q = id % g.Np;
  uint32_t pos = atomicAdd(&cursor[rw * g.H + (size_t)((d < 0 ? -d : d) - 1)], 1u);
  sorted[rw * g.Np + pos] = ((uint32_t)q << 1) | (d < 0 ? 1u : 0u);
}
// thread -> (rw, b): Bucket = sum of the segment
template <class F>
__global__ void __launch_bounds__(128) k_pip_accumulate(const Aff<F>* __restrict__ pts, const uint32_t* __restrict__ off,
                                                        const uint32_t* __restrict__ sorted, Jac<F>* __restrict__ buckets, size_t nrw,
                                                        pip_geom g) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nrw * g.H) return;
  size_t rw = id / g.H, b = id % g.H;
  uint32_t i0 = off[rw * (g.H + 1) + b], i1 = off[rw * (g.H + 1) + b + 1];
  Jac<F> acc;
  acc.set_inf();
  for (uint32_t i = i0; i < i1; i++) {
    uint32_t e = sorted[rw * g.Np + i];
    Aff<F> P = pts[e >> 1];
    if (e & 1) F::neg(P.y, P.y);
    Jac<F>::add_mixed(acc, acc, P);
  }
  buckets[id] = acc;
}
// thread -> (rw, grp): Q = sum_{k < G} (grp*G + k + 1) * Bucket[grp*G + k]
template <class F>
__global__ void __launch_bounds__(128) k_pip_bucket_reduce(const Jac<F>* __restrict__ buckets, Jac<F>* __restrict__ Q, size_t nrw, pip_geom g) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nrw * g.NG) return;
  size_t rw = id / g.NG;
  int grp = (int)(id % g.NG);
  const Jac<F>* bk = buckets + rw * g.H + (size_t)grp * g.G;
  Jac<F> run, acc;
  run.set_inf();
  acc.set_inf();
  for (int k = g.G - 1; k >= 0; k--) {
    Jac<F> t = bk[k];
    Jac<F>::add(run, run, t);
    Jac<F>::add(acc, acc, run);
  }
  // + (grp * G) * run   (double-and-add over the <= 12 bits of the group offset)
  uint32_t offs = (uint32_t)grp * (uint32_t)g.G;
  if (offs != 0 && !run.is_inf()) {
    Jac<F> t;
    t.set_inf();
    for (int bit = 31 - __clz(offs); bit >= 0; bit--) {
      Jac<F>::dbl(t, t);
      if ((offs >> bit) & 1) Jac<F>::add(t, t, run);
    }
    Jac<F>::add(acc, acc, t);
  }
  Q[id] = acc;
}
// thread -> rw = (row, w): Wd[rw] = 2^(c w) * Q[rw * NG]
template <class F>
__global__ void __launch_bounds__(64) k_pip_window_shift(const Jac<F>* __restrict__ Q, Jac<F>* __restrict__ Wd, size_t nrw, pip_geom g) {
  size_t rw = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (rw >= nrw) return;
  int w = (int)(rw % g.W);
  Jac<F> p = Q[rw * g.NG];
  if (!p.is_inf())
    for (int i = 0; i < g.c * w; i++) Jac<F>::dbl(p, p);
  Wd[rw] = p;
}


// ------------------------------------------------------------------ many equations, ONE scalar vector (C4 proofs)
// A multi-equation statement proves every equation under the SAME commitment randomness: the constant part of a proof
// element is  sum_t s[row][t] * B_e[t]  with one scalar vector for all equations e and per-equation bases (B_e = the
// equation's constants).  The bucket structure depends on the scalars only, so it is built ONCE (k_cs_digits, k_cs_sort:
// which sub-term goes to which bucket of which window) and every equation sums its own points along those lists:
//   k_cs_accumulate  thread (e, row, window, bucket): the bucket's points, endomorphism images taken on the fly
//   k_cs_reduce      thread (e, row, window): sum_b (b+1) Bucket_b by running sums
//   k_cs_combine     thread (e, row): Horner over the windows  ->  term slot 0 of (e, row)
// ~W * Np mixed additions per (equation, row) -- 13 per 64-bit sub-scalar at c = 5 -- against a 4-bit-window scalar
// multiplication (~64 doublings + 23 additions per sub-scalar) per term.
struct cs_geom {
  int c, W, H, rows;
  size_t N, Np, nt;  // constant terms, sub-terms (N * PARTS), terms per row of `sv`
};
// thread -> (row, t): signed c-bit digits of the PARTS sub-scalars of sv[row][t]:  digits[(row*W + w)*Np + t*PARTS + j]
template <class F>
__global__ void __launch_bounds__(128) k_cs_digits(const fr* __restrict__ sv, int16_t* __restrict__ digits, cs_geom g) {
  constexpr int PARTS = PipSplit<F>::PARTS;
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= g.N * g.rows) return;
  const size_t t = id % g.N;
  const int row = (int)(id / g.N);
  uint32_t k[8], sub[PARTS][4];
  fr_from_mont(k, sv[(size_t)row * g.nt + t]);
  PipSplit<F>::split(sub, k);
  const uint32_t half = 1u << (g.c - 1);
  for (int j = 0; j < PARTS; j++) {
    uint32_t carry = 0;
    for (int w = 0; w < g.W; w++) {
      uint32_t d = pip_bits(sub[j], w * g.c, g.c) + carry;
      int sd;
      if (d > half) {
        sd = (int)d - (int)(1u << g.c);
        carry = 1;
      } else {
        sd = (int)d;
        carry = 0;
      }
      digits[((size_t)row * g.W + w) * g.Np + t * PARTS + j] = (int16_t)sd;
    }
  }
}
// block -> rw = (row, window): counting sort of its sub-terms by |digit| (one thread: a few hundred items, once per call)
//   off[rw*(H+1) + b] = start of bucket b (b = |digit| - 1), off[.. + H] = number of non-zero digits
//   list[rw*Np + pos] = (q << 1) | negative
template <class F>
__global__ void k_cs_sort(const int16_t* __restrict__ digits, uint32_t* __restrict__ off, uint32_t* __restrict__ cursor,
                          uint32_t* __restrict__ list, cs_geom g) {
  if (threadIdx.x != 0) return;
  const size_t rw = blockIdx.x;
  const int16_t* d = digits + rw * g.Np;
  uint32_t* o = off + rw * (g.H + 1);
  uint32_t* cu = cursor + rw * g.H;
  uint32_t* l = list + rw * g.Np;
  for (int b = 0; b < g.H; b++) cu[b] = 0;
  for (size_t q = 0; q < g.Np; q++)
    if (d[q] != 0) cu[(d[q] < 0 ? -d[q] : d[q]) - 1]++;
  uint32_t run = 0;
  for (int b = 0; b < g.H; b++) {
    const uint32_t cnt = cu[b];
    o[b] = run;
    cu[b] = run;
    run += cnt;
  }
  o[g.H] = run;
  for (size_t q = 0; q < g.Np; q++)
    if (d[q] != 0) l[cu[(d[q] < 0 ? -d[q] : d[q]) - 1]++] = ((uint32_t)q << 1) | (d[q] < 0 ? 1u : 0u);
}
// thread -> (e, rw, b): buckets[(e*RW + rw)*H + b] = sum of the bucket's sub-terms of equation e
template <class F>
__global__ void __launch_bounds__(128) k_cs_accumulate(const Aff<F>* __restrict__ bases, const uint32_t* __restrict__ off,
                                                       const uint32_t* __restrict__ list, Jac<F>* __restrict__ buckets, size_t neq,
                                                       cs_geom g) {
  constexpr int PARTS = PipSplit<F>::PARTS;
  const size_t RW = (size_t)g.rows * g.W;
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= neq * RW * g.H) return;
  const size_t b = id % g.H, rw = (id / g.H) % RW, e = id / (g.H * RW);
  const uint32_t i0 = off[rw * (g.H + 1) + b], i1 = off[rw * (g.H + 1) + b + 1];
  Jac<F> acc;
  acc.set_inf();
  for (uint32_t i = i0; i < i1; i++) {
    const uint32_t en = list[rw * g.Np + i];
    const uint32_t q = en >> 1;
    Aff<F> P;
    PipSplit<F>::image(P, bases[e * g.N + q / PARTS], (int)(q % PARTS));
    if (en & 1) F::neg(P.y, P.y);
    Jac<F>::add_mixed(acc, acc, P);
  }
  buckets[id] = acc;
}
// thread -> (e, rw): S[e*RW + rw] = sum_b (b + 1) * Bucket_b   (running sums from the top bucket down)
template <class F>
__global__ void __launch_bounds__(128) k_cs_reduce(const Jac<F>* __restrict__ buckets, Jac<F>* __restrict__ S, size_t neq, cs_geom g) {
  const size_t RW = (size_t)g.rows * g.W;
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= neq * RW) return;
  const Jac<F>* bk = buckets + id * g.H;
  Jac<F> run, acc;
  run.set_inf();
  acc.set_inf();
  for (int b = g.H - 1; b >= 0; b--) {
    Jac<F> t = bk[b];
    Jac<F>::add(run, run, t);
    Jac<F>::add(acc, acc, run);
  }
  S[id] = acc;
}
// thread -> (e, row): out[(e*rows + row) * ostride] = sum_w 2^(c w) S[e][row][w]
template <class F>
__global__ void __launch_bounds__(64) k_cs_combine(const Jac<F>* __restrict__ S, Jac<F>* __restrict__ out, size_t ostride, size_t neq,
                                                   cs_geom g) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= neq * g.rows) return;
  const Jac<F>* s = S + id * g.W;
  Jac<F> acc = s[g.W - 1];
  for (int w = g.W - 2; w >= 0; w--) {
    for (int i = 0; i < g.c; i++) Jac<F>::dbl(acc, acc);
    Jac<F> t = s[w];
    Jac<F>::add(acc, acc, t);
  }
  out[id * ostride] = acc;
}

}  // namespace gs

namespace gsi {

// Window geometry.  The sub-scalars are < BOUND (G1: x^2 ~ 2^127.4, G2: |x| ~ 2^63.7), so the TOP window only takes
// top_vals = (BOUND >> ((W-1) c)) + 1 values: its buckets are fuller than the others by 2^(c-1) / top_vals, and a bucket is
// one thread's serial chain.  W windows must also cover bits + 1 so that the signed recoding never carries out of the top
// window (a carry window would put half of all points into ONE bucket).  c is chosen to minimise
//     max(longest chain, additions / lanes)   with chain = Np / min(H, top_vals), additions = W * Np.
// short_bits > 0: every scalar is known to be < 2^short_bits <= 2^63 (the weights of gs_verify_batch_rand), below the bound
// of the first sub-scalar on either group: the split leaves it where it is and the other sub-scalars are zero, so the
// windows only span short_bits.
template <class F>
inline pip_geom pip_choose(size_t N, int c_override, int short_bits = 0) {
  pip_geom g;
  g.parts = PipSplit<F>::PARTS;
  g.bits = short_bits > 0 ? short_bits : PipSplit<F>::BITS;
  g.N = N;
  g.Np = N * g.parts;
  // top 64 bits of the bound, for top_vals: G1 x^2 = 0xac45a4010001a402_0000000100000000, G2 |x| = 0xd201000000010000
  const unsigned long long bound_hi = short_bits > 0 ? ~0ull : (g.bits == 128 ? 0xac45a4010001a402ull : 0xd201000000010000ull);
  int best_c = 0;
  double best_t = 0;
  for (int c = 3; c <= 13; c++) {
    if (c_override >= 3 && c_override <= 13 && c != c_override) continue;
    const int W = (g.bits + 1 + c - 1) / c;
    const int top_lo = (W - 1) * c;                 // first bit of the top window
    const int top_bits = g.bits - top_lo;           // 1 .. c-? real bits in it
    if (top_bits < 1) continue;
    const double top_vals = (double)(bound_hi >> (64 - top_bits)) + 1.0;
    const double H = (double)(1 << (c - 1));
    const double chain = (double)g.Np / (top_vals < H ? top_vals : H) + 40.0;   // + the bucket reduction's own chain
    const double work = (double)W * ((double)g.Np + 3.0 * H) / 40000.0;          // additions / resident lanes
    const double t = chain > work ? chain : work;
    if (best_c == 0 || t < best_t) {
      best_c = c;
      best_t = t;
    }
  }
  g.c = best_c;
  g.H = 1 << (g.c - 1);
  g.W = (g.bits + 1 + g.c - 1) / g.c;
  g.G = g.H < 16 ? g.H : 16;
  g.NG = g.H / g.G;
  return g;
}

// terms[(e*rows + row) * ostride] = sum_{t < N} sv[row*nt + t] * bases[e*N + t] for e < neq: ONE scalar vector (the first
// equation's rows of `sv`) for all equations.  `terms` slots other than the written ones are untouched.
template <class F>
int shared_scalar_sums(gs_ctx* ctx, Scratch& sc, const fr* sv, size_t nt, int rows, const Aff<F>* bases, size_t N, size_t neq,
                       Jac<F>* terms, size_t ostride) {
  cs_geom g;
  g.rows = rows;
  g.N = N;
  g.Np = N * PipSplit<F>::PARTS;
  g.nt = nt;
  int lg = 0;
  while (((size_t)1 << (lg + 1)) <= g.Np) lg++;
  g.c = lg - 3 < 4 ? 4 : (lg - 3 > 10 ? 10 : lg - 3);  // ~Np / 8 points per bucket chain: W * (Np + 2 H) near its minimum
  g.H = 1 << (g.c - 1);
  g.W = (PipSplit<F>::BITS + 1 + g.c - 1) / g.c;
  const size_t RW = (size_t)rows * g.W;
  int16_t* digits;
  uint32_t *off, *cursor, *list;
  Jac<F>*buckets, *S;
  CUDA_TRY(sc.alloc(&digits, RW * g.Np));
  CUDA_TRY(sc.alloc(&off, RW * (g.H + 1)));
  CUDA_TRY(sc.alloc(&cursor, RW * g.H));
  CUDA_TRY(sc.alloc(&list, RW * g.Np));
  CUDA_TRY(sc.alloc(&buckets, neq * RW * g.H));
  CUDA_TRY(sc.alloc(&S, neq * RW));
  LAUNCH((k_cs_digits<F>), N * rows, sv, digits, g);
  LAUNCH_CFG((k_cs_sort<F>), RW * 32, 32, 0, digits, off, cursor, list, g);
  LAUNCH((k_cs_accumulate<F>), neq * RW * g.H, bases, off, list, buckets, neq, g);
  LAUNCH((k_cs_reduce<F>), neq * RW, buckets, S, neq, g);
  LAUNCH_CFG((k_cs_combine<F>), neq * rows, 64, 0, S, terms, ostride, neq, g);
  return GS_OK;
}

// out_rows[row * W] (Jacobian, row stride W) = sum_t sv[row][t] * (b0 | b1)[t]; returns the stride through *stride
template <class F>
int pippenger_rows(gs_ctx* ctx, Scratch& sc, const fr* sv, int rows, const Aff<F>* b0, size_t n0, const Aff<F>* b1, size_t n1,
                   Jac<F>** out_rows, size_t* stride, int c_override, int short_bits = 0) {
  const pip_geom g = pip_choose<F>(n0 + n1, c_override, short_bits);
  const size_t nrw = (size_t)rows * g.W;
  Aff<F>* pts;
  int16_t* digits;
  uint32_t *counts, *off, *cursor, *sorted;
  Jac<F>*buckets, *Q, *Wd;
  CUDA_TRY(sc.alloc(&pts, g.Np));
  CUDA_TRY(sc.alloc(&digits, nrw * g.Np));
  CUDA_TRY(sc.alloc(&counts, nrw * g.H));
  CUDA_TRY(sc.alloc(&off, nrw * (g.H + 1)));
  CUDA_TRY(sc.alloc(&cursor, nrw * g.H));
  CUDA_TRY(sc.alloc(&sorted, nrw * g.Np));
  CUDA_TRY(sc.alloc(&buckets, nrw * g.H));
  CUDA_TRY(sc.alloc(&Q, nrw * g.NG));
  CUDA_TRY(sc.alloc(&Wd, nrw));
  CUDA_TRY(cudaMemsetAsync(counts, 0, nrw * g.H * sizeof(uint32_t), ctx->stream));
  LAUNCH((k_pip_expand<F>), g.N, b0, n0, b1, n1, sv, rows, pts, digits, g);
  LAUNCH((k_pip_hist<F>), nrw * g.Np, digits, counts, nrw * g.Np, g);
  LAUNCH_CFG((k_pip_scan<F>), nrw * 256, 256, 0, counts, off, cursor, g);
  LAUNCH((k_pip_scatter<F>), nrw * g.Np, digits, cursor, sorted, nrw * g.Np, g);
  LAUNCH((k_pip_accumulate<F>), nrw * g.H, pts, off, sorted, buckets, nrw, g);
  LAUNCH((k_pip_bucket_reduce<F>), nrw * g.NG, buckets, Q, nrw, g);
  int rc = reduce_rows<F>(ctx, Q, (size_t)g.NG, (size_t)g.NG, (int)nrw, 1);
  if (rc) return rc;
  LAUNCH_CFG((k_pip_window_shift<F>), nrw, 64, 0, Q, Wd, nrw, g);
  rc = reduce_rows<F>(ctx, Wd, (size_t)g.W, (size_t)g.W, rows, 1);
  if (rc) return rc;
  *out_rows = Wd;
  *stride = (size_t)g.W;
  return GS_OK;
}

}  // namespace gsi
