"""Golden fixtures (tests/golden/golden.json, made by tests/golden/make_golden.py from the big-int oracle).
CPU: the oracle and the C restatement still reproduce them (drift guard).  GPU (-m gpu): the CUDA path
reproduces them byte for byte through the C ABI."""
import json
import os

import pytest

from gsutil import *  # noqa: F401,F403
from oracle import gs as ogs

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
H = bytes.fromhex


def test_oracle_reproduces_golden_pairing_and_crs():
    assert fp12_b(pairing(G1_GEN, G2_GEN_FP2)).hex() == GOLD["pairing_generators"]
    crs, _ = make_crs(1)
    assert crs_bytes(crs).hex() == GOLD["crs_seed1"]


def test_oracle_reproduces_one_golden_proof():
    crs, _ = make_crs(1)
    ty = 3                                     # Quad: the cheapest for the big-int oracle
    rng = SeededRng(700 + ty)
    c = GOLD["prove_verify"][ty]
    equ, xv, yv = random_instance(ty, c["m"], c["n"], crs, rng, zero_frac=0.25)
    xr, yr, T = draw_rands(ty, c["m"], c["n"], rng)
    proof = ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)
    assert [a.hex() for a in proof_bytes(ty, equ, proof)] == c["arrays"]


@pytest.fixture(scope="module")
def eng():
    import groth_sahai_rs_b200 as gsb
    e = gsb.Engine(0)
    e.crs_load(H(GOLD["crs_seed1"]))
    return e


@pytest.mark.gpu
def test_gpu_reproduces_golden(eng):
    assert eng.pairing(g1_b(G1_GEN), g2_b(G2_GEN_FP2)).hex() == GOLD["pairing_generators"]
    d = H(GOLD["crs_seed1_draws"])
    got = eng.crs_generate(d[:96], d[96:288], d[288:320], d[320:352], d[352:384], d[384:416])
    assert got.hex() == GOLD["crs_seed1"]
    ps = GOLD["pairing_sum"]
    assert eng.comt_pairing_sum(H(ps["xs"]), H(ps["ys"])).hex() == ps["comt"]
    for c in GOLD["prove_verify"]:
        ty, m, n = c["type"], c["m"], c["n"]
        a = [H(x) for x in c["arrays"]]
        if ty in (0, 1):
            assert eng.batch_commit_g1(H(c["xvars"]), H(c["xrand"])).hex() == c["arrays"][4]
        else:
            assert eng.batch_commit_scalar_b1(H(c["xvars"]), H(c["xrand"])).hex() == c["arrays"][4]
        if ty in (0, 2):
            assert eng.batch_commit_g2(H(c["yvars"]), H(c["yrand"])).hex() == c["arrays"][5]
        else:
            assert eng.batch_commit_scalar_b2(H(c["yvars"]), H(c["yrand"])).hex() == c["arrays"][5]
        pi, th = eng.prove(ty, m, n, a[0], a[1], a[2], H(c["xvars"]), H(c["yvars"]), H(c["xrand"]), H(c["yrand"]), H(c["T"]))
        assert pi.hex() == c["arrays"][6] and th.hex() == c["arrays"][7]
        assert eng.verify(ty, m, n, *a) is True
        bad = list(a)
        bad[7] = bad[7][96:192] + bad[7][:96] + bad[7][192:]      # swap the two coordinates of theta[0]
        assert eng.verify(ty, m, n, *bad) is False
