// Final exponentiation kernel (arkworks exponent: easy part, then (x-1)^2 (x+p)(x^2+p^2-1) + 3) and the
// comparison against the expected ComT entry.  Reference: ark-ec final_exponentiation as reached from
// src/data_structures.rs:484-502.
#include "ctx.h"
#include "pairing.cuh"

using namespace gs;

namespace gs {

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


}  // namespace gs

int gsi::launch_final_exp(gs_ctx* ctx, const fp12* F, size_t nprob, int nchunk, fp12* out_comt, uint8_t* ok4, const fp12* target) {
  LAUNCH(k_final_exp, nprob * 4, F, nprob, nchunk, out_comt, ok4, target);
  return GS_OK;
}
