"""The reference's own unit / integration tests, replayed through the host mirror of its API (api.py) on the GPU:
RNG draw order, batch == singles, commit_and_prove == commit x2 + prove, append, equ_type tags, iota_T
commutativity, verify's panics.  Each test names the reference test it mirrors (paths under /root/reference)."""
import pytest

from gsutil import *  # noqa: F401,F403
from oracle import gs as ogs

pytestmark = pytest.mark.gpu


class ReplayRng:
    """Replayable host randomness in ABI bytes (the role ark_std::test_rng() plays with DETERMINISTIC_TEST_RNG=1)."""

    def __init__(self, seed):
        self.s = SeededRng(seed)
        self.log = []

    def fr(self):
        v = self.s.fr()
        self.log.append(v)
        return fr_b(v)

    def g1(self):
        p = self.s.g1()
        self.log.append(p)
        return g1_b(p)

    def g2(self):
        q = self.s.g2()
        self.log.append(q)
        return g2_b(q)


@pytest.fixture(scope="module")
def env():
    import groth_sahai_rs_b200 as gsb
    from groth_sahai_rs_b200 import api
    eng = gsb.Engine(0)
    rng = ReplayRng(1)
    crs = api.CRS.generate_crs(rng, eng)
    return eng, api, crs, rng


def test_crs_draw_order_and_binding_key(env):
    """generator.rs:137-182: gt_gen == e(g1, g2); u2 = t1 u1, v2 = t2 v1 by replaying the RNG (p1, p2, a1, a2, t1, t2)."""
    eng, api, crs, rng = env
    p1, p2, a1, a2, t1, t2 = rng.log[:6]
    assert crs.g1_gen == g1_b(p1) and crs.g2_gen == g2_b(p2)
    assert crs.gt_gen == fp12_b(pairing(p1, p2))
    q1, q2 = g1_mul(p1, a1), g2_mul(p2, a2)
    assert com1_i(crs.u[0]) == (p1, q1) and com2_i(crs.v[0]) == (p2, q2)
    assert com1_i(crs.u[1]) == (g1_mul(p1, t1), g1_mul(q1, t1))            # binding: u[1].1 == t1 * q1
    assert com2_i(crs.v[1]) == (g2_mul(p2, t2), g2_mul(q2, t2))
    assert crs.gt_gen != fp12_b(FP12_ONE)                                   # non-degeneracy


def test_hiding_crs_differs_only_in_the_last_key_entries():
    """generator.rs:62-77 (prepare_simulated_hinding_key, dead code upstream): same draws as the binding key; only
    u[1].1 = t1*q1 - g1 and v[1].1 = t2*q2 - g2 change, so W1 = u[1] + iota_1(g1) = t1*u[0] (scalar commitments
    carry no information) while gt_gen, the generators and u[0], v[0] are untouched."""
    import groth_sahai_rs_b200 as gsb
    from groth_sahai_rs_b200 import api
    eng = gsb.Engine(0)
    r1, r2 = ReplayRng(44), ReplayRng(44)
    real, hid = api.CRS.generate_crs(r1, eng), api.CRS.generate_hiding_crs(r2, eng)
    assert r1.log == r2.log and len(r2.log) == 6
    p1, p2, a1, a2, t1, t2 = r2.log
    assert (hid.u[0], hid.v[0], hid.g1_gen, hid.g2_gen, hid.gt_gen) == (real.u[0], real.v[0], real.g1_gen, real.g2_gen, real.gt_gen)
    q1, q2 = g1_mul(p1, a1), g2_mul(p2, a2)
    assert com1_i(hid.u[1]) == (g1_mul(p1, t1), G1.add(g1_mul(q1, t1), G1.neg(p1)))
    assert com2_i(hid.v[1]) == (g2_mul(p2, t2), G2.add(g2_mul(q2, t2), G2.neg(p2)))
    assert hid.u[1] != real.u[1] and hid.v[1] != real.v[1]
    # W1 under the hiding key is t1 * u[0]: iota_1'(x) = x*W1 lies in the span of u[0]
    w1 = api.Com1.scalar_linear_map(fr_b(1), hid)
    assert com1_i(w1) == (g1_mul(p1, t1), g1_mul(q1, t1))
    # proofs made under the hiding key still verify under it (completeness does not depend on the key kind)
    crs_o = ogs.generate_crs(p1, p2, a1, a2, t1, t2)
    for ty in (0, 3):
        _, equ, xvars, yvars, _, _ = _equation(api, crs_o, ty, 60 + ty)
        cp = equ.commit_and_prove(xvars, yvars, hid, ReplayRng(5))
        assert equ.verify(cp, hid) is True


def test_batch_commit_equals_sequence_of_singles(env):
    """commit.rs:439-548: batch_commit_* == the singles appended, under the same RNG stream (row-major (r1, r2))."""
    eng, api, crs, _ = env
    src = SeededRng(5)
    xs, ys = [g1_b(src.g1()) for _ in range(3)], [g2_b(src.g2()) for _ in range(3)]
    sx, sy = [fr_b(src.fr()) for _ in range(3)], [fr_b(src.fr()) for _ in range(3)]
    for batch, single, vals in ((api.batch_commit_G1, api.commit_G1, xs), (api.batch_commit_G2, api.commit_G2, ys),
                                (api.batch_commit_scalar_to_B1, api.commit_scalar_to_B1, sx),
                                (api.batch_commit_scalar_to_B2, api.commit_scalar_to_B2, sy)):
        b = batch(vals, crs, ReplayRng(9))
        r2 = ReplayRng(9)
        acc = single(vals[0], crs, r2)
        for v in vals[1:]:
            acc.append(single(v, crs, r2))                                  # Commit::append, commit.rs:42-51
        assert acc.coms == b.coms and acc.rand == b.rand
        assert len(b.coms) == 3 and len(b.rand) == 3


def _equation(api, crs_o, ty, seed):
    rng = SeededRng(seed)
    equ, xv, yv = random_instance(ty, 2, 1, crs_o, rng, zero_frac=0.3)
    cls = [api.PPE, api.MSMEG1, api.MSMEG2, api.QuadEqu][ty]
    e = cls([x for x in _split_enc(enc_A(ty, equ.a_consts), ty, "A")], [x for x in _split_enc(enc_B(ty, equ.b_consts), ty, "B")],
            [[fr_b(g) for g in row] for row in equ.gamma], enc_T(ty, equ.target))
    return equ, e, _split_enc(enc_A(ty, xv), ty, "A"), _split_enc(enc_B(ty, yv), ty, "B"), xv, yv


def _split_enc(b, ty, side):
    size = (96 if ty in (0, 1) else 32) if side == "A" else (192 if ty in (0, 2) else 32)
    return [b[i:i + size] for i in range(0, len(b), size)]


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_commit_and_prove_equals_commits_then_prove(env, ty):
    """prove.rs:537-589, 652-702, 763-813, 879-932: commit_and_prove == batch_commit x2 then prove under a replayed
    RNG (draw order xcoms -> ycoms -> T); prove.rs:510-535: the equ_type tag; tests/prover.rs: the proof verifies."""
    eng, api, crs, rng = env
    crs_o = ogs.generate_crs(*rng.log[:6])
    equ_o, equ, xvars, yvars, xv_o, yv_o = _equation(api, crs_o, ty, 30 + ty)
    cp = equ.commit_and_prove(xvars, yvars, crs, ReplayRng(77))
    r2 = ReplayRng(77)
    xc = (api.batch_commit_G1 if ty in (0, 1) else api.batch_commit_scalar_to_B1)(xvars, crs, r2)
    yc = (api.batch_commit_G2 if ty in (0, 2) else api.batch_commit_scalar_to_B2)(yvars, crs, r2)
    ep = equ.prove(xvars, yvars, xc, yc, crs, r2)
    assert (cp.xcoms.coms, cp.xcoms.rand, cp.ycoms.coms, cp.ycoms.rand) == (xc.coms, xc.rand, yc.coms, yc.rand)
    assert (cp.equ_proofs[0].pi, cp.equ_proofs[0].theta, cp.equ_proofs[0].rand) == (ep.pi, ep.theta, ep.rand)
    assert cp.equ_proofs[0].equ_type == ty == equ.get_type()
    assert equ.verify(cp, crs) is True
    # the oracle, fed the same draws, produces the same bytes (cross-implementation vector the reference lacks)
    xr = [[fr_i(x) for x in row] for row in xc.rand]
    yr = [[fr_i(x) for x in row] for row in yc.rand]
    T = [[fr_i(x) for x in row] for row in ep.rand]
    want = ogs.commit_and_prove(equ_o, xv_o, yv_o, crs_o, xr, yr, T)
    assert cp.xcoms.coms == [com1_b(c) for c in want.xcoms.coms] and cp.ycoms.coms == [com2_b(c) for c in want.ycoms.coms]
    assert ep.pi == [com2_b(c) for c in want.equ_proofs[0].pi] and ep.theta == [com1_b(c) for c in want.equ_proofs[0].theta]


def test_verify_panics_like_the_reference(env):
    """verifier.rs:25-26, 59-60: exactly one EquProof and a matching equ_type, else panic (AssertionError here)."""
    eng, api, crs, rng = env
    crs_o = ogs.generate_crs(*rng.log[:6])
    _, equ, xvars, yvars, _, _ = _equation(api, crs_o, 0, 41)
    cp = equ.commit_and_prove(xvars, yvars, crs, ReplayRng(3))
    two = api.CProof(cp.xcoms, cp.ycoms, cp.equ_proofs * 2)
    with pytest.raises(AssertionError):
        equ.verify(two, crs)
    wrong = api.CProof(cp.xcoms, cp.ycoms, [api.EquProof(cp.equ_proofs[0].pi, cp.equ_proofs[0].theta, 1, cp.equ_proofs[0].rand)])
    with pytest.raises(AssertionError):
        equ.verify(wrong, crs)
    # a tampered proof is rejected (the negative test the reference lacks)
    bad = api.CProof(cp.xcoms, cp.ycoms, [api.EquProof(cp.equ_proofs[0].pi[::-1], cp.equ_proofs[0].theta, 0, cp.equ_proofs[0].rand)])
    assert equ.verify(bad, crs) is False


def test_verify_batch_randomized_flag(env):
    """api.verify_batch(randomized=True) -- the opt-in of SURVEY.md 8f.4: one randomised check of the batch, the exact
    per-proof verification only when it rejects; the answers equal the exact call's, honest or not."""
    eng, api, crs, rng = env
    crs_o = ogs.generate_crs(*rng.log[:6])
    eqs, proofs = [], []
    for i in range(4):
        _, equ, xvars, yvars, _, _ = _equation(api, crs_o, 0, 60 + i)
        eqs.append(equ)
        proofs.append(equ.commit_and_prove(xvars, yvars, crs, ReplayRng(90 + i)))
    assert api.verify_batch(eqs, proofs, crs, randomized=True) == [True] * 4
    ep = proofs[2].equ_proofs[0]
    bad = list(proofs)
    bad[2] = api.CProof(proofs[2].xcoms, proofs[2].ycoms, [api.EquProof(ep.pi[::-1], ep.theta, 0, ep.rand)])
    assert api.verify_batch(eqs, bad, crs, randomized=True) == [True, True, False, True] == api.verify_batch(eqs, bad, crs)


def test_iota_t_commutes_with_the_maps(env):
    """tests/commit.rs:22-85: iota_T(f(x, y)) == F(iota_1(x), iota_2(y)) for the four equation types."""
    eng, api, crs, rng = env
    src = SeededRng(51)
    x, y, sx, sy = src.g1(), src.g2(), src.fr(), src.fr()
    P = lambda a, b: api.ComT.pairing(a, b, eng)
    assert api.ComT.linear_map_PPE(fp12_b(pairing(x, y)), crs) == P(api.Com1.linear_map(g1_b(x)), api.Com2.linear_map(g2_b(y)))
    assert api.ComT.linear_map_MSMEG1(g1_b(g1_mul(x, sy)), crs) == P(api.Com1.linear_map(g1_b(x)),
                                                                     api.Com2.scalar_linear_map(fr_b(sy), crs))
    assert api.ComT.linear_map_MSMEG2(g2_b(g2_mul(y, sx)), crs) == P(api.Com1.scalar_linear_map(fr_b(sx), crs),
                                                                     api.Com2.linear_map(g2_b(y)))
    assert api.ComT.linear_map_quad(fr_b(sx * sy % R), crs) == P(api.Com1.scalar_linear_map(fr_b(sx), crs),
                                                                 api.Com2.scalar_linear_map(fr_b(sy), crs))
    # batch maps == element-wise maps (data_structures.rs:1467-1512)
    xs = [g1_b(src.g1()) for _ in range(3)]
    assert api.Com1.batch_linear_map(xs) == [api.Com1.linear_map(v) for v in xs]
    ss = [fr_b(src.fr()) for _ in range(3)]
    assert api.Com1.batch_scalar_linear_map(ss, crs) == [api.Com1.scalar_linear_map(v, crs) for v in ss]
    assert api.Com2.batch_scalar_linear_map(ss, crs) == [api.Com2.scalar_linear_map(v, crs) for v in ss]


def test_group_axioms_on_com_types(env):
    """data_structures.rs:943-1265: add-zero, commutativity, Sum, neg, sub, scalar_mul == component-wise mul."""
    eng, api, crs, rng = env
    src = SeededRng(61)
    a, b = com1_b((src.g1(), src.g1())), com1_b((src.g1(), src.g1()))
    assert api.Com1.add(a, api.Com1.zero(), eng) == a
    assert api.Com1.add(a, b, eng) == api.Com1.add(b, a, eng)
    assert api.Com1.sub(a, a, eng) == api.Com1.zero()
    assert api.Com1.add(a, api.Com1.neg(a, eng), eng) == api.Com1.zero()
    assert api.Com1.sum([a, b, a], eng) == api.Com1.add(api.Com1.add(a, b, eng), a, eng)
    s = src.fr()
    pa = com1_i(a)
    assert com1_i(api.Com1.scalar_mul(a, fr_b(s), eng)) == (g1_mul(pa[0], s), g1_mul(pa[1], s))
    c, d = com2_b((src.g2(), src.g2())), com2_b((src.g2(), src.g2()))
    assert api.Com2.add(c, d, eng) == api.Com2.add(d, c, eng) and api.Com2.sub(c, c, eng) == api.Com2.zero()
    t = api.ComT.pairing(a, c, eng)
    assert api.ComT.add(t, api.ComT.zero(eng), eng) == t
    assert api.ComT.sub(t, t, eng) == api.ComT.zero(eng)
    assert api.ComT.as_matrix(t)[1][0] == t[1152:1728]                       # row-major entries (:1361-1377)
