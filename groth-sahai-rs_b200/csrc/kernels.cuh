// CUDA kernels of the Groth-Sahai hot path (sm_100a).  See DESIGN.md for the data layout.
//
// "Slot" layout used by every pairing-product evaluation (ComT::pairing, pairing_sum,
// verify): a problem is K pairs (X_k in Com1, Y_k in Com2), and the wanted ComT is
//     ComT[a][b] = prod_k e(X_k.a, Y_k.b)          (src/data_structures.rs:494-502)
// Points are stored SoA over problems so that a warp (32 consecutive problems, same slot,
// same coordinate) reads contiguous memory:
//     X[(a*K + k) * nprob + p]   g1_aff        Y[(b*K + k) * nprob + p]   g2_aff
//     L[((b*K + k) * 68 + step) * nprob + p]   line_coeffs  (output of k_g2_prepare)
#pragma once
#include "pairing.cuh"

namespace gs {

struct crs_dev {  // device copy of the key + derived constants  (generator.rs:36-42)
  g1_aff u[2][2];      // u[k][a]
  g2_aff v[2][2];      // v[k][b]
  g1_aff w1[2];        // W1 = u2 + (O, g1)       data_structures.rs:325
  g2_aff w2[2];        // W2 = v2 + (O, g2)       data_structures.rs:370
  g1_aff neg_u[2][2];  // -u[k][a]
  g1_aff neg_w1[2];
  g1_aff g1;
  g2_aff g2;
};

struct verify_shape {  // slot bookkeeping shared by host and device
  int type, m, n;
  int groupA, groupB;  // 1: constants are group elements (iota), 0: scalars (iota')
  int cx, cy;
  int sB, nB, sPi, sTh, sT, K;
  int n_out;    // MSM outputs per problem: n (+1 scalar-B) (+1 Quad target)
  int nbases;   // m (+1 when A is scalar: W1 is an extra base)
  int nchunk;   // MSM base chunks
};
constexpr int GS_MSM_CHUNK = 16;

inline verify_shape make_verify_shape(int type, int m, int n) {
  verify_shape s;
  s.type = type;
  s.m = m;
  s.n = n;
  s.groupA = (type == 0 || type == 1);
  s.groupB = (type == 0 || type == 2);
  s.cx = s.groupA ? 2 : 1;  // x-variables (and A) are G1 for PPE / MSMEG1  => R is m x 2, |pi| = 2
  s.cy = s.groupB ? 2 : 1;  // y-variables (and B) are G2 for PPE / MSMEG2  => S is n x 2, |theta| = 2
  s.sB = n;
  s.nB = s.groupB ? m : 1;
  s.sPi = s.sB + s.nB;
  s.sTh = s.sPi + s.cx;
  s.sT = s.sTh + s.cy;
  s.K = s.sT + (type == 0 ? 0 : 1);
  s.n_out = n + (s.groupB ? 0 : 1) + (type == 3 ? 1 : 0);
  s.nbases = m + (s.groupA ? 0 : 1);
  s.nchunk = (s.nbases + GS_MSM_CHUNK - 1) / GS_MSM_CHUNK;
  return s;
}

// ------------------------------------------------------------------ G2 preparation
// one thread per G2 point q = (b*K + k) * nprob + p
__global__ void __launch_bounds__(128) k_g2_prepare(const g2_aff* __restrict__ Y, line_coeffs* __restrict__ L,
                                                    uint8_t* __restrict__ yinf, size_t npoints, size_t nprob) {
  size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= npoints) return;
  g2_aff pt = Y[q];
  bool inf = pt.is_inf();
  yinf[q] = inf ? 1 : 0;
  if (inf) return;
  size_t bk = q / nprob, p = q % nprob;
  g2_prepare(L + (bk * GS_NUM_LINES) * nprob + p, nprob, pt);
}

// ------------------------------------------------------------------ Miller accumulation
// thread -> (p, e = 2a+b, chunk);  F[(chunk*4 + e) * nprob + p] = conj( prod over its slots )
__global__ void __launch_bounds__(128) k_miller(const g1_aff* __restrict__ X, const uint8_t* __restrict__ yinf,
                                                const line_coeffs* __restrict__ L, fp12* __restrict__ F,
                                                size_t nprob, int K, int S, int nchunk) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nprob * 4 * (size_t)nchunk) return;
  size_t p = id % nprob;
  int e = (int)((id / nprob) & 3);
  int ch = (int)(id / (nprob * 4));
  int a = e >> 1, b = e & 1;
  int k0 = ch * S, k1 = min(K, k0 + S);
  fp12 f;
  f.set_one();
  // anything to do at all?
  bool any = false;
  for (int k = k0; k < k1; k++) {
    if (yinf[((size_t)b * K + k) * nprob + p]) continue;
    if (X[((size_t)a * K + k) * nprob + p].is_inf()) continue;
    any = true;
  }
  if (any) {
    int idx = 0;
    for (int bit = 62; bit >= 0; bit--) {
      fp12::sqr(f, f);
      int nl = ((GS_X_ABS >> bit) & 1) ? 2 : 1;
      for (int t = 0; t < nl; t++, idx++) {
        for (int k = k0; k < k1; k++) {
          size_t bk = (size_t)b * K + k;
          if (yinf[bk * nprob + p]) continue;
          const g1_aff* P = &X[((size_t)a * K + k) * nprob + p];
          fp px = P->x, py = P->y;
          if (px.is_zero() && py.is_zero()) continue;
          line_coeffs l = L[(bk * GS_NUM_LINES + idx) * nprob + p];
          miller_apply_line(f, l, px, py);
        }
      }
    }
    fp12::conj(f, f);
  }
  F[((size_t)ch * 4 + e) * nprob + p] = f;
}

// ------------------------------------------------------------------ final exponentiation (+ compare)
// thread -> (p, e).  f = prod_chunks F;  g = FE(f).
//   out_comt != null : out_comt[p].e[e] = g
//   ok != null       : ok[e*nprob + p] = (g == expected), expected = target[p] for PPE entry 3, else 1
__global__ void __launch_bounds__(128) k_final_exp(const fp12* __restrict__ F, size_t nprob, int nchunk,
                                                   fp12* __restrict__ out_comt, uint8_t* __restrict__ ok,
                                                   const fp12* __restrict__ target) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nprob * 4) return;
  size_t p = id % nprob;
  int e = (int)(id / nprob);
  fp12 f = F[(size_t)e * nprob + p];
  for (int ch = 1; ch < nchunk; ch++) {
    fp12 g = F[((size_t)ch * 4 + e) * nprob + p];
    fp12::mul(f, f, g);
  }
  fp12 one;
  one.set_one();
  fp12 g;
  if (f.equals(one)) {
    g = one;
  } else {
    final_exponentiation(g, f);
  }
  if (out_comt) out_comt[p * 4 + e] = g;
  if (ok) {
    bool good;
    if (target != nullptr && e == 3) {
      fp12 t = target[p];
      good = g.equals(t);
    } else {
      good = g.equals(one);
    }
    ok[(size_t)e * nprob + p] = good ? 1 : 0;
  }
}

__global__ void k_and4(const uint8_t* __restrict__ ok4, uint8_t* __restrict__ out, size_t nprob) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nprob) return;
  out[p] = ok4[p] & ok4[nprob + p] & ok4[2 * nprob + p] & ok4[3 * nprob + p];
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

// ------------------------------------------------------------------ verify: slot assembly
struct verify_args {
  const void* a_consts;  // [p][n]  g1_aff | fr
  const void* b_consts;  // [p][m]  g2_aff | fr
  const fr* gamma;       // [p][m][n]
  const void* target;    // [p]     fp12 | g1_aff | g2_aff | fr
  const g1_aff* xcoms;   // [p][m][2]
  const g2_aff* ycoms;   // [p][n][2]
  const g2_aff* pi;      // [p][cx][2]
  const g1_aff* theta;   // [p][cy][2]
};

// thread -> (p, k): fills every slot that is a plain copy / negation.  X slots [0,n), the scalar-B
// slot and the Quad target slot are written later by k_vmsm_reduce.
__global__ void k_verify_assemble(verify_shape s, verify_args v, const crs_dev* __restrict__ crs, g1_aff* __restrict__ X,
                                  g2_aff* __restrict__ Y, size_t nprob) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nprob * (size_t)s.K) return;
  size_t p = id % nprob;
  int k = (int)(id / nprob);
  const int K = s.K;
  g1_aff x0, x1;
  g2_aff y0, y1;
  bool writeX = true;
  if (k < s.n) {  // (P_j, d_j)
    writeX = false;
    y0 = v.ycoms[(p * s.n + k) * 2 + 0];
    y1 = v.ycoms[(p * s.n + k) * 2 + 1];
  } else if (k < s.sPi) {
    if (s.groupB) {  // (c_i, (O, B_i))
      int i = k - s.sB;
      x0 = v.xcoms[(p * s.m + i) * 2 + 0];
      x1 = v.xcoms[(p * s.m + i) * 2 + 1];
      y0.set_inf();
      y1 = ((const g2_aff*)v.b_consts)[p * s.m + i];
    } else {  // (sum_i b_i c_i, W2)
      writeX = false;
      y0 = crs->w2[0];
      y1 = crs->w2[1];
    }
  } else if (k < s.sTh) {  // (-u_k, pi_k)
    int j = k - s.sPi;
    x0 = crs->neg_u[j][0];
    x1 = crs->neg_u[j][1];
    y0 = v.pi[(p * s.cx + j) * 2 + 0];
    y1 = v.pi[(p * s.cx + j) * 2 + 1];
  } else if (k < s.sT) {  // (-theta_k, v_k)
    int j = k - s.sTh;
    x0 = v.theta[(p * s.cy + j) * 2 + 0];
    x1 = v.theta[(p * s.cy + j) * 2 + 1];
    fp::neg(x0.y, x0.y);
    fp::neg(x1.y, x1.y);
    y0 = crs->v[j][0];
    y1 = crs->v[j][1];
  } else {  // target slot
    if (s.type == 1) {  // (-(O, t), W2)
      x0.set_inf();
      x1 = ((const g1_aff*)v.target)[p];
      fp::neg(x1.y, x1.y);
      y0 = crs->w2[0];
      y1 = crs->w2[1];
    } else if (s.type == 2) {  // (-W1, (O, t))
      x0 = crs->neg_w1[0];
      x1 = crs->neg_w1[1];
      y0.set_inf();
      y1 = ((const g2_aff*)v.target)[p];
    } else {  // Quad: (-(t W1), W2)
      writeX = false;
      y0 = crs->w2[0];
      y1 = crs->w2[1];
    }
  }
  if (writeX) {
    X[((size_t)0 * K + k) * nprob + p] = x0;
    X[((size_t)1 * K + k) * nprob + p] = x1;
  }
  Y[((size_t)0 * K + k) * nprob + p] = y0;
  Y[((size_t)1 * K + k) * nprob + p] = y1;
}

// ------------------------------------------------------------------ verify: G1-side statement MSM
// P_j = iota(A_j) + sum_i Gamma_ij c_i      (re-association of verifier.rs:39-42, SURVEY.md §8a ‡)
// thread -> (p, jj, a, chunk): Straus (shared doublings) over <= GS_MSM_CHUNK bases
__global__ void __launch_bounds__(128) k_vmsm_partial(verify_shape s, verify_args v, const crs_dev* __restrict__ crs,
                                                      g1_jac* __restrict__ part, size_t nprob) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = nprob * (size_t)s.n_out * 2 * s.nchunk;
  if (id >= total) return;
  size_t p = id % nprob;
  size_t r = id / nprob;
  int a = (int)(r & 1);
  r >>= 1;
  int jj = (int)(r % s.n_out);
  int ch = (int)(r / s.n_out);

  uint32_t sc[GS_MSM_CHUNK][8];
  g1_aff base[GS_MSM_CHUNK];
  int cnt = 0;
  int i0 = ch * GS_MSM_CHUNK, i1 = min(s.nbases, i0 + GS_MSM_CHUNK);
  for (int i = i0; i < i1; i++) {
    fr sv;
    bool have = false;
    g1_aff bp;
    if (jj < s.n) {
      if (i < s.m) {
        sv = v.gamma[(p * s.m + i) * s.n + jj];
        bp = v.xcoms[(p * s.m + i) * 2 + a];
        have = true;
      } else {  // scalar A: extra base W1 with scalar a_j
        sv = ((const fr*)v.a_consts)[p * s.n + jj];
        bp = crs->w1[a];
        have = true;
      }
    } else if (jj == s.n && !s.groupB) {  // C_B = sum_i b_i c_i
      if (i < s.m) {
        sv = ((const fr*)v.b_consts)[p * s.m + i];
        bp = v.xcoms[(p * s.m + i) * 2 + a];
        have = true;
      }
    } else {  // Quad target: t * W1
      if (i == 0) {
        sv = ((const fr*)v.target)[p];
        bp = crs->w1[a];
        have = true;
      }
    }
    if (!have || bp.is_inf() || sv.is_zero()) continue;
    fr_from_mont(sc[cnt], sv);
    base[cnt] = bp;
    cnt++;
  }
  g1_jac acc;
  acc.set_inf();
  if (cnt > 0) {
    // highest set bit over the chunk
    int top = -1;
    for (int w = 7; w >= 0 && top < 0; w--) {
      uint32_t o = 0;
      for (int i = 0; i < cnt; i++) o |= sc[i][w];
      if (o) top = w * 32 + 31 - __clz(o);
    }
    for (int bit = top; bit >= 0; bit--) {
      g1_jac::dbl(acc, acc);
      for (int i = 0; i < cnt; i++)
        if ((sc[i][bit >> 5] >> (bit & 31)) & 1) g1_jac::add_mixed(acc, acc, base[i]);
    }
  }
  part[(((size_t)ch * s.n_out + jj) * 2 + a) * nprob + p] = acc;
}

// thread -> (p, jj, a): sum the chunk partials, add iota_1(A_j), negate the Quad target, normalise, write slot
__global__ void __launch_bounds__(128) k_vmsm_reduce(verify_shape s, verify_args v, const g1_jac* __restrict__ part,
                                                     g1_aff* __restrict__ X, size_t nprob) {
  size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= nprob * (size_t)s.n_out * 2) return;
  size_t p = id % nprob;
  size_t r = id / nprob;
  int a = (int)(r & 1);
  int jj = (int)(r >> 1);
  g1_jac acc = part[((size_t)jj * 2 + a) * nprob + p];
  for (int ch = 1; ch < s.nchunk; ch++) {
    g1_jac t = part[(((size_t)ch * s.n_out + jj) * 2 + a) * nprob + p];
    g1_jac::add(acc, acc, t);
  }
  int slot;
  if (jj < s.n) {
    slot = jj;
    if (s.groupA && a == 1) {
      g1_aff A = ((const g1_aff*)v.a_consts)[p * s.n + jj];
      g1_jac::add_mixed(acc, acc, A);
    }
  } else if (jj == s.n && !s.groupB) {
    slot = s.sB;
  } else {
    slot = s.sT;
    g1_jac::neg(acc, acc);
  }
  g1_aff out;
  g1_jac::to_affine(out, acc);
  X[((size_t)a * s.K + slot) * nprob + p] = out;
}

}  // namespace gs
