"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the oracle on
the same seeded inputs.  Bit-exact (integer / byte work): every comparison is equality of
canonical values.  Run on the B200 box with `pytest -m gpu`."""
import pytest

from gsutil import *  # noqa: F401,F403
from oracle import gs as ogs
from oracle.bls12_381 import multi_pairing

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import groth_sahai_rs_b200 as gsb
    return gsb.Engine(0)


@pytest.fixture(scope="module")
def crs_pair(eng):
    crs, draws = make_crs(1)
    eng.crs_load(crs_bytes(crs))
    return crs, draws


def test_pairing_matches_oracle(eng):
    rng = SeededRng(2)
    ps = [G1_GEN, rng.g1(), rng.g1(), None, rng.g1()]
    qs = [G2_GEN_FP2, rng.g2(), rng.g2(), rng.g2(), None]
    out = eng.pairing(b"".join(g1_b(p) for p in ps), b"".join(g2_b(q) for q in qs))
    for i, (p, q) in enumerate(zip(ps, qs)):
        assert fp12_i(out[576 * i:576 * (i + 1)]) == pairing(p, q), i
    # bilinearity on the device alone
    a, b = rng.fr(), rng.fr()
    o2 = eng.pairing(g1_b(g1_mul(G1_GEN, a)) + g1_b(g1_mul(G1_GEN, a * b % R)), g2_b(g2_mul(G2_GEN_FP2, b)) + g2_b(G2_GEN_FP2))
    assert o2[:576] == o2[576:]


def test_comt_pairing_and_sum(eng):
    rng = SeededRng(3)
    xs = [(rng.g1(), rng.g1()), (None, rng.g1()), (rng.g1(), None)]
    ys = [(rng.g2(), rng.g2()), (rng.g2(), rng.g2()), (None, rng.g2())]
    xb, yb = b"".join(com1_b(x) for x in xs), b"".join(com2_b(y) for y in ys)
    out = eng.comt_pairing(xb, yb)
    for i in range(3):
        assert comt_i(out[2304 * i:2304 * (i + 1)]) == ogs.comt_pairing(xs[i], ys[i]), i
    # entry order [e(x0,y0), e(x0,y1), e(x1,y0), e(x1,y1)]  (data_structures.rs:1361-1377)
    assert fp12_i(out[576:1152]) == pairing(xs[0][0], ys[0][1])
    s = eng.comt_pairing_sum(xb, yb)
    assert comt_i(s) == ogs.comt_pairing_sum(xs, ys)
    # pairing_sum == sum of pairings (data_structures.rs:1381-1407)
    acc = ogs.comt_zero()
    for i in range(3):
        acc = ogs.comt_add(acc, comt_i(out[2304 * i:2304 * (i + 1)]))
    assert comt_i(s) == acc
    # identity inputs give the GT identity (data_structures.rs:1313-1343); empty sum = zero
    z = eng.comt_pairing(com1_b((None, None)), com2_b(ys[0]))
    assert comt_i(z) == [FP12_ONE] * 4
    assert comt_i(eng.comt_pairing_sum(b"", b"")) == [FP12_ONE] * 4


def test_pairing_sum_many_pairs_chunked(eng):
    """k large enough that the Miller kernel splits the slots over several threads."""
    rng = SeededRng(4)
    k = 9
    dl_x = [(rng.fr(), rng.fr()) for _ in range(k)]
    dl_y = [(rng.fr(), rng.fr()) for _ in range(k)]
    xs = [(g1_mul(G1_GEN, a), g1_mul(G1_GEN, b)) for a, b in dl_x]
    ys = [(g2_mul(G2_GEN_FP2, a), g2_mul(G2_GEN_FP2, b)) for a, b in dl_y]
    s = comt_i(eng.comt_pairing_sum(b"".join(com1_b(x) for x in xs), b"".join(com2_b(y) for y in ys)))
    e = pairing(G1_GEN, G2_GEN_FP2)
    for a in range(2):
        for b in range(2):
            ex = sum(dl_x[i][a] * dl_y[i][b] for i in range(k)) % R
            assert s[2 * a + b] == e.pow(ex)


def test_crs_generate(eng):
    crs, (p1, p2, a1, a2, t1, t2) = make_crs(7)
    out = eng.crs_generate(g1_b(p1), g2_b(p2), fr_b(a1), fr_b(a2), fr_b(t1), fr_b(t2))
    assert out == crs_bytes(crs)
    # gt_gen == e(g1, g2) and binding-key structure (generator.rs:137-182)
    assert fp12_i(out[-576:]) == pairing(p1, p2)
    assert crs.u[1][1] == g1_mul(crs.u[0][1], t1)


def test_linear_maps(eng, crs_pair):
    crs, _ = crs_pair
    rng = SeededRng(5)
    t1, t2, tf = rng.g1(), rng.g2(), rng.fr()
    gt = pairing(rng.g1(), rng.g2())
    assert comt_i(eng.comt_linear_map(0, fp12_b(gt))) == ogs.comt_linear_map_ppe(gt)
    assert comt_i(eng.comt_linear_map(1, g1_b(t1))) == ogs.comt_linear_map_msmeg1(t1, crs)
    assert comt_i(eng.comt_linear_map(2, g2_b(t2))) == ogs.comt_linear_map_msmeg2(t2, crs)
    assert comt_i(eng.comt_linear_map(3, fr_b(tf))) == ogs.comt_linear_map_quad(tf, crs)


def test_batch_commits(eng, crs_pair):
    crs, _ = crs_pair
    rng = SeededRng(6)
    xv = [rng.g1(), None, crs.g1_gen, g1_mul(crs.g1_gen, 2), rng.g1()]
    R2 = [[rng.fr(), rng.fr()], [rng.fr(), 0], [0, 0], [1, R - 1], [rng.fr(), rng.fr()]]
    got = eng.batch_commit_g1(b"".join(g1_b(x) for x in xv), frmat_b(R2))
    exp = ogs.batch_commit_g1(xv, crs, R2)
    assert got == b"".join(com1_b(c) for c in exp.coms)
    yv = [rng.g2(), None, crs.g2_gen, rng.g2()]
    S2 = [[rng.fr(), rng.fr()], [0, rng.fr()], [0, 0], [rng.fr(), rng.fr()]]
    got = eng.batch_commit_g2(b"".join(g2_b(y) for y in yv), frmat_b(S2))
    exp = ogs.batch_commit_g2(yv, crs, S2)
    assert got == b"".join(com2_b(c) for c in exp.coms)
    xs = [rng.fr(), 0, 1, rng.fr()]
    r1 = [[rng.fr()], [rng.fr()], [0], [rng.fr()]]
    got = eng.batch_commit_scalar_b1(frs_b(xs), frmat_b(r1))
    assert got == b"".join(com1_b(c) for c in ogs.batch_commit_scalar_to_b1(xs, crs, r1).coms)
    got = eng.batch_commit_scalar_b2(frs_b(xs), frmat_b(r1))
    assert got == b"".join(com2_b(c) for c in ogs.batch_commit_scalar_to_b2(xs, crs, r1).coms)
    assert eng.batch_commit_g1(b"", b"") == b""


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_prove_and_verify_reference_scenario(eng, crs_pair, ty):
    """tests/prover.rs:24-172: 2 x-vars, 1 y-var, Gamma = [[5],[0]], B = [O, c2]; prove bit-exact, verify accepts."""
    crs, _ = crs_pair
    rng = SeededRng(10 + ty)
    g1, g2 = crs.g1_gen, crs.g2_gen
    gamma = [[5], [0]]
    xv = [g1_mul(g1, 2), g1_mul(g1, 3)] if ty in (0, 1) else [2, 3]
    yv = [g2_mul(g2, 4)] if ty in (0, 2) else [4]
    ca, cb = rng.fr(), rng.fr()
    a = [g1_mul(g1, ca)] if ty in (0, 1) else [ca]
    b = [None, g2_mul(g2, cb)] if ty in (0, 2) else [0, cb]
    val = (3 * cb + ca * 4 + 2 * 4 * 5) % R
    target = {0: lambda: crs.gt_gen.pow(val), 1: lambda: g1_mul(g1, val), 2: lambda: g2_mul(g2, val), 3: lambda: val}[ty]()
    equ = ogs.Equation(ty, a, b, gamma, target)
    xr, yr, T = draw_rands(ty, 2, 1, rng)
    exp = ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)
    assert ogs.verify(equ, exp, crs)
    pi, th = eng.prove(ty, 2, 1, enc_A(ty, a), enc_B(ty, b), frmat_b(gamma), enc_A(ty, xv), enc_B(ty, yv),
                       frmat_b(xr), frmat_b(yr), frmat_b(T))
    ep = exp.equ_proofs[0]
    assert pi == b"".join(com2_b(c) for c in ep.pi)
    assert th == b"".join(com1_b(c) for c in ep.theta)
    arrs = proof_bytes(ty, equ, exp)
    assert eng.verify(ty, 2, 1, *arrs) is True
    # tampered target / tampered proof element must be rejected (SURVEY.md §4 gap)
    bad_t = {0: lambda: target * crs.gt_gen, 1: lambda: G1.add(target, g1), 2: lambda: G2.add(target, g2),
             3: lambda: (target + 1) % R}[ty]()
    arrs2 = list(arrs)
    arrs2[3] = enc_T(ty, bad_t)
    assert eng.verify(ty, 2, 1, *arrs2) is False
    arrs3 = list(arrs)
    arrs3[7] = com1_b((ep.theta[0][0], G1.add(ep.theta[0][1], g1))) + arrs[7][192:]
    assert eng.verify(ty, 2, 1, *arrs3) is False


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_prove_verify_random_shapes(eng, crs_pair, ty):
    crs, _ = crs_pair
    rng = SeededRng(20 + ty)
    m, n = 3, 2
    equ, xv, yv = random_instance(ty, m, n, crs, rng, zero_frac=0.2)
    xr, yr, T = draw_rands(ty, m, n, rng)
    exp = ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)
    pi, th = eng.prove(ty, m, n, enc_A(ty, equ.a_consts), enc_B(ty, equ.b_consts), frmat_b(equ.gamma), enc_A(ty, xv),
                       enc_B(ty, yv), frmat_b(xr), frmat_b(yr), frmat_b(T))
    ep = exp.equ_proofs[0]
    assert pi == b"".join(com2_b(c) for c in ep.pi)
    assert th == b"".join(com1_b(c) for c in ep.theta)
    assert eng.verify(ty, m, n, *proof_bytes(ty, equ, exp)) is True


def test_verify_batch_mixed_verdicts(eng, crs_pair):
    """Batch of PPE proofs (4x4, the C1/C5 shape) with tampered ones at known positions."""
    crs, _ = crs_pair
    rng = SeededRng(30)
    m = n = 4
    insts = []
    for i in range(3):
        equ, xv, yv = random_instance(0, m, n, crs, rng, gamma_small=(i == 1))
        xr, yr, T = draw_rands(0, m, n, rng)
        insts.append((equ, ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)))
    assert ogs.verify(*insts[0], crs)
    rows = [proof_bytes(0, e, p) for e, p in insts]
    # replicate to 70 proofs (more than two warps), tamper a few
    count = 70
    tampered = {5, 33, 69}
    cols = [[] for _ in range(8)]
    for i in range(count):
        r = list(rows[i % 3])
        if i in tampered:
            if i % 2:
                r[3] = fp12_b(insts[i % 3][0].target * crs.gt_gen)
            else:  # swap pi[0] and pi[1]
                r[6] = r[6][384:] + r[6][:384]
        for c in range(8):
            cols[c].append(r[c])
    ok = eng.verify_batch(0, count, m, n, *[b"".join(c) for c in cols])
    assert list(ok) == [0 if i in tampered else 1 for i in range(count)]


def test_matmul_kats(eng):
    """The reference's only literal KATs (data_structures.rs:1678-1947)."""
    f = lambda m: frmat_b(m)
    out = eng.fr_matmul(1, 3, 1, f([[1, 2, 3]]), f([[4], [5], [6]]))
    assert fr_i(out) == 32
    a = [[1, 2, 3], [4, 5, 6]]
    b = [[7, 8, 9, 10], [11, 12, 13, 14], [15, 16, 17, 18]]
    out = eng.fr_matmul(2, 3, 4, f(a), f(b))
    got = [[fr_i(out[(i * 4 + j) * 32:(i * 4 + j + 1) * 32]) for j in range(4)] for i in range(2)]
    assert got == [[74, 80, 86, 92], [173, 188, 203, 218]]
    assert got == ogs.fr_right_mul(a, b)
    # "in the exponent" Com1 / Com2 versions (:1951-2006): lhs (2x3) * col-vector of Com
    rng = SeededRng(31)
    col1 = [[(rng.g1(), rng.g1())], [(None, rng.g1())], [(rng.g1(), rng.g1())]]
    lhs = [[1, 2, 3], [rng.fr(), 0, rng.fr()]]
    out = eng.com1_matmul(2, 3, 1, f(lhs), b"".join(com1_b(r[0]) for r in col1))
    exp = ogs.com_left_mul(col1, lhs, 1)
    assert out == b"".join(com1_b(r[0]) for r in exp)
    col2 = [[(rng.g2(), rng.g2())], [(rng.g2(), None)], [(rng.g2(), rng.g2())]]
    out = eng.com2_matmul(2, 3, 1, f(lhs), b"".join(com2_b(r[0]) for r in col2))
    exp = ogs.com_left_mul(col2, lhs, 2)
    assert out == b"".join(com2_b(r[0]) for r in exp)


def test_error_behaviour(eng, crs_pair):
    import groth_sahai_rs_b200 as gsb
    with pytest.raises(gsb.GsError):   # empty variable list: the reference panics
        eng.verify_batch(0, 1, 0, 0, b"", b"", b"", bytes(576), b"", b"", bytes(768), bytes(384))
    with pytest.raises(gsb.GsError):   # pairing_sum length mismatch (data_structures.rs:495)
        eng.comt_pairing_sum(bytes(192), b"")


# ---------------------------------------------------------------- commitment-group arithmetic
def test_com_group_ops_match_oracle(eng):
    """Com1 / Com2 Add, Sub, Neg, Sum (impl_base_commit_groups!, data_structures.rs:162-255) incl. the
    exceptional cases: identity operands, P + P, P + (-P)."""
    rng = SeededRng(21)
    p, q = rng.g1(), rng.g1()
    xs = [(p, q), (None, q), (p, None), (None, None), (p, q), (rng.g1(), rng.g1())]
    ys = [(q, p), (p, None), (None, q), (None, None), (p, ogs.G1.neg(q)), (rng.g1(), rng.g1())]
    xb, yb = b"".join(com1_b(x) for x in xs), b"".join(com1_b(y) for y in ys)
    add = eng.elementwise("com1", "add", xb, yb)
    sub = eng.elementwise("com1", "sub", xb, yb)
    neg = eng.elementwise("com1", "neg", xb)
    for i, (x, y) in enumerate(zip(xs, ys)):
        assert com1_i(add[192 * i:192 * (i + 1)]) == ogs.com1_add(x, y), i
        assert com1_i(sub[192 * i:192 * (i + 1)]) == ogs.com1_add(x, ogs.com1_neg(y)), i
        assert com1_i(neg[192 * i:192 * (i + 1)]) == ogs.com1_neg(x), i
    acc = (None, None)
    for x in xs:
        acc = ogs.com1_add(acc, x)
    assert com1_i(eng.group_sum("com1", xb)) == acc
    assert eng.group_sum("com1", b"") == bytes(192)                      # Sum of nothing = Com1::zero()
    u, v = rng.g2(), rng.g2()
    as_ = [(u, v), (None, v), (u, u), (rng.g2(), rng.g2()), (rng.g2(), None)]
    bs_ = [(v, u), (u, None), (u, ogs.G2.neg(u)), (rng.g2(), rng.g2()), (None, None)]
    ab, bb = b"".join(com2_b(x) for x in as_), b"".join(com2_b(y) for y in bs_)
    add, sub, neg = eng.elementwise("com2", "add", ab, bb), eng.elementwise("com2", "sub", ab, bb), eng.elementwise("com2", "neg", ab)
    for i, (x, y) in enumerate(zip(as_, bs_)):
        assert com2_i(add[384 * i:384 * (i + 1)]) == ogs.com2_add(x, y), i
        assert com2_i(sub[384 * i:384 * (i + 1)]) == ogs.com2_add(x, ogs.com2_neg(y)), i
        assert com2_i(neg[384 * i:384 * (i + 1)]) == ogs.com2_neg(x), i
    acc = (None, None)
    for x in as_:
        acc = ogs.com2_add(acc, x)
    assert com2_i(eng.group_sum("com2", ab)) == acc


def test_comt_group_ops_match_oracle(eng):
    """ComT Add (GT product), Sub, Neg (conjugate), Sum, Zero (data_structures.rs:391-479)."""
    rng = SeededRng(22)
    xs = [(rng.g1(), rng.g1()) for _ in range(3)]
    ys = [(rng.g2(), rng.g2()) for _ in range(3)]
    t = eng.comt_pairing(b"".join(com1_b(x) for x in xs), b"".join(com2_b(y) for y in ys))
    ts = [comt_i(t[2304 * i:2304 * (i + 1)]) for i in range(3)]
    a, b = t[:2 * 2304], t[2304:]
    add, sub, neg = eng.elementwise("comt", "add", a, b), eng.elementwise("comt", "sub", a, b), eng.elementwise("comt", "neg", a)
    for i in range(2):
        assert comt_i(add[2304 * i:2304 * (i + 1)]) == ogs.comt_add(ts[i], ts[i + 1])
        assert comt_i(sub[2304 * i:2304 * (i + 1)]) == ogs.comt_add(ts[i], ogs.comt_neg(ts[i + 1]))
        assert comt_i(neg[2304 * i:2304 * (i + 1)]) == ogs.comt_neg(ts[i])
    acc = ogs.comt_zero()
    for x in ts:
        acc = ogs.comt_add(acc, x)
    assert comt_i(eng.group_sum("comt", t)) == acc
    # sum of pairings == pairing_sum (data_structures.rs:1381-1407), now entirely on the device
    assert eng.group_sum("comt", t) == eng.comt_pairing_sum(b"".join(com1_b(x) for x in xs), b"".join(com2_b(y) for y in ys))
    assert comt_i(eng.group_sum("comt", b"")) == ogs.comt_zero()
    # x - x = zero
    assert comt_i(eng.elementwise("comt", "sub", a, a)[:2304]) == ogs.comt_zero()


def test_mat_trait_mirror(eng, crs_pair):
    """The element-wise and structural parts of the Mat trait (data_structures.rs:588-643, 768-823) through
    api.py: add, neg, scalar_mul, transpose, right_mul for Matrix<Fr> and Matrix<Com1|Com2>."""
    from groth_sahai_rs_b200 import api
    rng = SeededRng(23)
    A = [[rng.fr() for _ in range(3)] for _ in range(2)]
    B = [[rng.fr() for _ in range(3)] for _ in range(2)]
    enc = lambda m: [[fr_b(x) for x in r] for r in m]
    dec = lambda m: [[fr_i(x) for x in r] for r in m]
    assert dec(api.fr_add(enc(A), enc(B), eng)) == ogs.fr_add(A, B)
    assert dec(api.fr_neg(enc(A), eng)) == ogs.fr_neg(A)
    s = rng.fr()
    assert dec(api.fr_scalar_mul(enc(A), fr_b(s), eng)) == ogs.fr_scalar_mul(A, s)
    assert dec(api.mat_transpose(enc(A))) == ogs.fr_transpose(A)
    with pytest.raises(AssertionError):
        api.fr_add(enc(A), enc(ogs.fr_transpose(B)), eng)
    C = [[(rng.g1(), rng.g1()) for _ in range(2)] for _ in range(3)]          # 3 x 2 Com1
    D = [[(rng.g1(), None) for _ in range(2)] for _ in range(3)]
    e1 = lambda m: [[com1_b(x) for x in r] for r in m]
    d1 = lambda m: [[com1_i(x) for x in r] for r in m]
    assert d1(api.com_add(e1(C), e1(D), 1, eng)) == ogs.com_mat_add(C, D, 1)
    assert d1(api.com_neg(e1(C), 1, eng)) == [[ogs.com1_neg(x) for x in r] for r in C]
    assert d1(api.com_scalar_mul(e1(C), fr_b(s), 1, eng)) == [[ogs.com1_scalar_mul(x, s) for x in r] for r in C]
    rhs = [[rng.fr() for _ in range(4)] for _ in range(2)]                     # 2 x 4 scalars
    assert d1(api.com_right_mul(e1(C), enc(rhs), 1, eng)) == ogs.com_right_mul(C, rhs, 1)
    E2 = [[(rng.g2(), rng.g2())] for _ in range(2)]                            # 2 x 1 Com2
    e2 = lambda m: [[com2_b(x) for x in r] for r in m]
    d2 = lambda m: [[com2_i(x) for x in r] for r in m]
    lhs = [[rng.fr(), rng.fr()] for _ in range(3)]
    assert d2(api.com_left_mul(e2(E2), enc(lhs), 2, eng)) == ogs.com_left_mul(E2, lhs, 2)
    assert d2(api.com_scalar_mul(e2(E2), fr_b(s), 2, eng)) == [[ogs.com2_scalar_mul(x, s) for x in r] for r in E2]
    # the reference's scalar_linear_map through the mirror (W recomputed per element upstream, :325)
    crs, _ = crs_pair
    from groth_sahai_rs_b200.api import CRS as ApiCRS
    key = ApiCRS.from_bytes(crs_bytes(crs), eng)
    x = rng.fr()
    assert com1_i(api.Com1.scalar_linear_map(fr_b(x), key)) == ogs.com1_scalar_linear_map(x, crs)
    assert com2_i(api.Com2.scalar_linear_map(fr_b(x), key)) == ogs.com2_scalar_linear_map(x, crs)
