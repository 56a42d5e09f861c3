// Host/device shared plain-data descriptions: device copy of the CRS, verify slot bookkeeping.
#pragma once
#include "curve.cuh"

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
  int chunk, nchunk;  // MSM bases per thread (<= GS_MSM_CHUNK) and the number of such chunks
  int rank, world;  // statement sharding: slot k belongs to rank k % world; 0, 1 = everything
  // MSM sharding by BASE (gs_verify_sharded): this rank sums, for EVERY output, only the bases i = brank (mod bworld); the
  // partial sums are exchanged.  Gamma then holds only this rank's rows: gm = #{i < m : i = brank mod bworld} (m when 1).
  int brank, bworld, gm;
  // coordinates per base / per MSM output: 2 (Com1 = two G1 points, summed independently); 1 in gs_verify_batch_rand, where
  // the bases are the FOLDED commitments sigma c.0 + tau c.1 (verify_args::xfold) and every output is a single point
  int na;
  GS_HD int nb_own() const { return bworld <= 1 ? nbases : (nbases > brank ? (nbases - brank + bworld - 1) / bworld : 0); }
  GS_HD int base_at(int io) const { return bworld <= 1 ? io : brank + io * bworld; }
  GS_HD bool owns(int slot) const { return world <= 1 || slot % world == rank; }
  // slot that MSM output jj is written to: jj < n -> jj; the scalar-B sum -> sB; the Quad target -> sT
  GS_HD int out_slot(int jj) const { return jj < n ? jj : ((jj == n && !groupB) ? sB : sT); }
  // the MSM outputs this rank owns, enumerated densely (jo -> jj) so that every lane of a warp has work
  GS_HD int n_main_owned() const { return world <= 1 ? n : (n > rank ? (n - rank + world - 1) / world : 0); }
  GS_HD int n_out_owned() const {
    int c = n_main_owned();
    for (int jj = n; jj < n_out; jj++) c += owns(out_slot(jj)) ? 1 : 0;
    return c;
  }
  GS_HD int owned_out(int jo) const {
    const int nm = n_main_owned();
    if (jo < nm) return world <= 1 ? jo : rank + jo * world;
    jo -= nm;
    for (int jj = n; jj < n_out; jj++)
      if (owns(out_slot(jj))) {
        if (jo == 0) return jj;
        jo--;
      }
    return -1;
  }
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
  s.chunk = GS_MSM_CHUNK;
  s.nchunk = (s.nbases + GS_MSM_CHUNK - 1) / GS_MSM_CHUNK;
  s.rank = 0;
  s.world = 1;
  s.brank = 0;
  s.bworld = 1;
  s.gm = m;
  s.na = 2;
  return s;
}

inline void set_base_shard(verify_shape& s, int brank, int bworld) {
  s.brank = brank;
  s.bworld = bworld;
  s.gm = bworld <= 1 ? s.m : (s.m > brank ? (s.m - brank + bworld - 1) / bworld : 0);
}

inline void set_msm_chunk(verify_shape& s, int chunk) {
  s.chunk = chunk < 1 ? 1 : (chunk > GS_MSM_CHUNK ? GS_MSM_CHUNK : chunk);
  const int nb = s.nb_own() > 0 ? s.nb_own() : 1;
  s.nchunk = (nb + s.chunk - 1) / s.chunk;
}

struct verify_args {
  const void* a_consts;  // [p][n]  g1_aff | fr
  const void* b_consts;  // [p][m]  g2_aff | fr
  const fr* gamma;       // [p][m][n]
  const void* target;    // [p]     fp12 | g1_aff | g2_aff | fr
  const g1_aff* xcoms;   // [p][m][2]
  const g2_aff* ycoms;   // [p][n][2]
  const g2_aff* pi;      // [p][cx][2]
  const g1_aff* theta;   // [p][cy][2]
  // folded mode (verify_shape::na == 1): bases xfold[p][nbases], addends afold[p][n] = tau_p A_j (group-valued A); the MSM
  // outputs go to pair fold_map[slot] of the folded problem (fold_X1[p*fold_Kw + .]) or, for CRS slots (fold_map < 0), to
  // fold_Xfix[(-fold_map - 1)*nprob + p]
  const g1_aff* xfold = nullptr;
  const g1_aff* afold = nullptr;
  const int* fold_map = nullptr;
  int fold_Kw = 0;
  g1_aff* fold_X1 = nullptr;
  g1_aff* fold_Xfix = nullptr;
};

constexpr int GS_VTAB = 8;  // multiples 1B..8B per base in the verify-side Straus tables

}  // namespace gs
