// Block-wide Montgomery batch inversion and Jacobian -> affine normalisation (one Fp inversion per block).
#pragma once
#include "curve.cuh"

namespace gs {

// ------------------------------------------------------------------ block-level batch inversion
// Montgomery's trick as a product tree in shared memory with CONTIGUOUS active threads
// (3 products per element + one Fermat inversion per block).  All `NT` threads must call.
// z == 0 is passed through as 0.   sm: 2*NT fp.
template <int NT>
__device__ void block_batch_inv(fp& z, fp* sm) {
  int t = threadIdx.x;
  bool zero = z.is_zero();
  fp v = z;
  if (zero) fp_one(v);
  sm[NT + t] = v;
  __syncthreads();
  for (int half = NT / 2; half >= 1; half >>= 1) {
    if (t < half) {
      fp a = sm[2 * (half + t)], b = sm[2 * (half + t) + 1];
      fp::mul(a, a, b);
      sm[half + t] = a;
    }
    __syncthreads();
  }
  if (t == 0) {
    fp r = sm[1];
    fp_inv(r, r);
    sm[1] = r;
  }
  __syncthreads();
  for (int half = 1; half <= NT / 2; half <<= 1) {
    if (t < half) {
      int i = half + t;
      fp inv_i = sm[i], l = sm[2 * i], r = sm[2 * i + 1], nl, nr;
      fp::mul(nl, inv_i, r);
      fp::mul(nr, inv_i, l);
      sm[2 * i] = nl;
      sm[2 * i + 1] = nr;
    }
    __syncthreads();
  }
  z = sm[NT + t];
  if (zero) z.set_zero();
}

// Jacobian -> affine for a whole block at once (one field inversion per block).
template <int NT>
__device__ void block_to_affine(g1_aff& out, const g1_jac& p, fp* sm) {
  fp zi = p.Z;
  block_batch_inv<NT>(zi, sm);
  g1_jac::to_affine_with_zinv(out, p, zi);
}
template <int NT>
__device__ void block_to_affine(g2_aff& out, const g2_jac& p, fp* sm) {
  // 1/z = conj(z) / (z0^2 + z1^2): batch the Fp norm inversion
  fp n, t;
  fp::sqr(n, p.Z.c0);
  fp::sqr(t, p.Z.c1);
  fp::add(n, n, t);
  block_batch_inv<NT>(n, sm);
  fp2 zi;
  fp::mul(zi.c0, p.Z.c0, n);
  fp::mul(t, p.Z.c1, n);
  fp::neg(zi.c1, t);
  g2_jac::to_affine_with_zinv(out, p, zi);
}

}  // namespace gs
