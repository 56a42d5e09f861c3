#!/usr/bin/env python3
"""Regenerates tests/golden/golden.json from the big-int oracle (oracle/bls12_381.py, oracle/gs.py).

The reference holds no byte-level vectors for this path (parity unpinned against arkworks
bits, no Rust toolchain here), so these fixtures pin the ORACLE against drift and give the GPU tests fixed
byte strings to hit; the oracle itself is pinned by the public constants / identities of tests/test_oracle.py.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from gsutil import *  # noqa: E402,F401,F403
from oracle import gs as ogs  # noqa: E402


def build():
    out = {}
    out["pairing_generators"] = fp12_b(pairing(G1_GEN, G2_GEN_FP2)).hex()
    crs, draws = make_crs(1)
    out["crs_seed1"] = crs_bytes(crs).hex()
    out["crs_seed1_draws"] = (g1_b(draws[0]) + g2_b(draws[1]) + b"".join(fr_b(x) for x in draws[2:])).hex()
    cases = []
    for ty in range(4):
        rng = SeededRng(700 + ty)
        m, n = (2, 3) if ty % 2 else (3, 2)
        equ, xv, yv = random_instance(ty, m, n, crs, rng, zero_frac=0.25)
        xr, yr, T = draw_rands(ty, m, n, rng)
        proof = ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)
        assert ogs.verify(equ, proof, crs)
        arrs = proof_bytes(ty, equ, proof)
        cases.append({
            "type": ty, "m": m, "n": n,
            "xvars": enc_A(ty, xv).hex(), "yvars": enc_B(ty, yv).hex(),
            "xrand": frmat_b(xr).hex(), "yrand": frmat_b(yr).hex(), "T": frmat_b(T).hex(),
            "arrays": [a.hex() for a in arrs],          # a_consts b_consts gamma target xcoms ycoms pi theta
        })
    out["prove_verify"] = cases
    # ComT::pairing_sum of 3 (Com1, Com2) pairs, one of them with identity coordinates
    rng = SeededRng(800)
    xs = [(rng.g1(), rng.g1()), (None, rng.g1()), (rng.g1(), rng.g1())]
    ys = [(rng.g2(), rng.g2()), (rng.g2(), rng.g2()), (rng.g2(), None)]
    out["pairing_sum"] = {
        "xs": b"".join(com1_b(c) for c in xs).hex(), "ys": b"".join(com2_b(c) for c in ys).hex(),
        "comt": b"".join(fp12_b(e) for e in ogs.comt_pairing_sum(xs, ys)).hex(),
    }
    return out


if __name__ == "__main__":
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(build(), f, indent=1)
    print("wrote golden.json")
