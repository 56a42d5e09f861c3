"""GPU parity of the wire-format kernels (serial.cu) against oracle/serialize.py: the public generator
encodings, random points, identity, every rejection class of Validate::Yes, canonical Fr / GT bytes."""
import random

import pytest

from gsutil import *  # noqa: F401,F403
from oracle import serialize as ser
from oracle.bls12_381 import P, Fp2
from test_serialize import G1_GEN_COMPRESSED, G2_GEN_COMPRESSED

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import groth_sahai_rs_b200 as gsb
    return gsb.Engine(0)


def test_g1_compress_decompress(eng):
    rng = SeededRng(61)
    pts = [G1_GEN, G1.neg(G1_GEN), None] + [rng.g1() for _ in range(29)]
    wire = eng.serialize("g1", b"".join(g1_b(p) for p in pts))
    assert wire[:48] == G1_GEN_COMPRESSED
    assert wire == b"".join(ser.g1_compress(p) for p in pts)
    back, ok = eng.deserialize("g1", wire)
    assert ok == b"\x01" * len(pts)
    assert back == b"".join(g1_b(p) for p in pts)


def test_g2_compress_decompress(eng):
    rng = SeededRng(62)
    pts = [G2_GEN_FP2, G2.neg(G2_GEN_FP2), None] + [rng.g2() for _ in range(13)]
    wire = eng.serialize("g2", b"".join(g2_b(p) for p in pts))
    assert wire[:96] == G2_GEN_COMPRESSED
    assert wire == b"".join(ser.g2_compress(p) for p in pts)
    back, ok = eng.deserialize("g2", wire)
    assert ok == b"\x01" * len(pts)
    assert back == b"".join(g2_b(p) for p in pts)


def test_decompress_rejections_match_oracle(eng):
    """Every class Validate::Yes rejects: not compressed, x >= p, x off the curve, point outside the subgroup;
    plus random byte strings -- verdict and decoded point must equal the oracle's for each."""
    rnd = random.Random(63)
    cases = [bytes([G1_GEN_COMPRESSED[0] & 0x7F]) + G1_GEN_COMPRESSED[1:]]          # compression flag missing
    bad = bytearray(P.to_bytes(48, "big"))
    bad[0] |= 0x80
    cases.append(bytes(bad))                                                        # x = p
    cases.append(bytes([0xC0]) + bytes(47))                                         # infinity
    cases.append(bytes([0xE0]) + bytes(47))                                         # sort flag + infinity: rejected
    cases.append(bytes([0xC0]) + bytes([7]) * 47)                                   # infinity with non-zero x bytes: rejected
    cases.append(bytes([0xC1]) + bytes(47))                                         # infinity with a low bit of byte 0 set
    for x in range(2, 40):                                                          # small x: off-curve / off-subgroup
        e = bytearray(x.to_bytes(48, "big"))
        e[0] |= 0x80 | (0x20 if x % 2 else 0)
        cases.append(bytes(e))
    for _ in range(20):
        e = bytearray(rnd.randbytes(48))
        e[0] = (e[0] & 0x1F) | 0x80
        e[0] &= 0x99                                                                # keep x < p most of the time
        cases.append(bytes(e))
    back, ok = eng.deserialize("g1", b"".join(cases))
    seen = set()
    for i, c in enumerate(cases):
        want_ok, want_pt = ser.g1_decompress(c)
        assert bool(ok[i]) == want_ok, (i, c.hex())
        assert back[96 * i:96 * (i + 1)] == g1_b(want_pt if want_ok else None), i
        seen.add(want_ok)
    assert seen == {True, False}
    # without the subgroup check, curve points of the wrong order are returned (Validate::No + on-curve)
    x = 2
    while True:
        x += 1
        y = ser.fp_sqrt((x ** 3 + 4) % P)
        if y is not None and not ser.in_subgroup_g1((x, y)):
            break
    e = bytearray(x.to_bytes(48, "big"))
    e[0] |= 0x80
    b2, ok2 = eng.deserialize("g1", bytes(e), check_subgroup=False)
    assert ok2 == b"\x01" and g1_i(b2)[0] == x
    assert eng.deserialize("g1", bytes(e))[1] == b"\x00"
    # G2: off-subgroup curve point, c0 with stray top bits, x.c1 >= p
    cases2 = []
    c0 = 0
    while len(cases2) < 2:
        c0 += 1
        xx = Fp2(c0, 1)
        yy = ser.fp2_sqrt(xx * xx * xx + Fp2(4, 4))
        e = bytearray((1).to_bytes(48, "big") + c0.to_bytes(48, "big"))
        e[0] |= 0x80
        if yy is not None:
            cases2.append(bytes(e))
    e = bytearray(G2_GEN_COMPRESSED)
    e[48] |= 0x80
    cases2.append(bytes(e))
    e = bytearray(P.to_bytes(48, "big") + bytes(48))
    e[0] |= 0x80
    cases2.append(bytes(e))
    cases2.append(bytes([0xC0]) + bytes(95))                                        # canonical infinity
    cases2.append(bytes([0xE0]) + bytes(95))                                        # sort flag + infinity
    cases2.append(bytes([0xC0]) + bytes(70) + b"\x01" + bytes(24))                  # infinity with a non-zero x.c0 byte
    cases2.append(G2_GEN_COMPRESSED)
    back, ok = eng.deserialize("g2", b"".join(cases2))
    for i, c in enumerate(cases2):
        if c[48] & 0xE0:
            want_ok, want_pt = False, None            # 0x80 on the second coordinate: integer >= 2^383 > p
        else:
            want_ok, want_pt = ser.g2_decompress(c)
        assert bool(ok[i]) == want_ok, i
        assert back[192 * i:192 * (i + 1)] == g2_b(want_pt if want_ok else None), i


def test_fr_and_gt_bytes(eng):
    rng = SeededRng(64)
    xs = [0, 1, R - 1] + [rng.fr() for _ in range(20)]
    wire = eng.serialize("fr", frs_b(xs))
    assert wire == b"".join(ser.fr_to_bytes(x) for x in xs)
    back, ok = eng.deserialize("fr", wire + R.to_bytes(32, "little") + bytes([255]) * 32)
    assert ok == b"\x01" * len(xs) + b"\x00\x00"
    assert back[:32 * len(xs)] == frs_b(xs)
    g = eng.pairing(g1_b(G1_GEN) + g1_b(rng.g1()), g2_b(G2_GEN_FP2) + g2_b(rng.g2()))
    w = eng.serialize("gt", g)
    assert w == b"".join(ser.fp12_to_bytes(fp12_i(g[576 * i:576 * (i + 1)])) for i in range(2))
    back, ok = eng.deserialize("gt", w)
    assert ok == b"\x01\x01" and back == g
    bad = bytearray(w[:576])
    bad[5 * 48:6 * 48] = P.to_bytes(48, "little")
    assert eng.deserialize("gt", bytes(bad))[1] == b"\x00"
    # PairingOutput's Valid::check: f^r == 1.  Zero, a random Fp12 and a cyclotomic element of the wrong order are rejected
    from oracle.bls12_381 import Fp2, Fp6, Fp12, FP12_ONE
    rnd = Fp12(Fp6(*[Fp2(rng.r.randrange(P), rng.r.randrange(P)) for _ in range(3)]),
               Fp6(*[Fp2(rng.r.randrange(P), rng.r.randrange(P)) for _ in range(3)]))
    cyc = rnd.pow((P ** 6 - 1) * (P ** 2 + 1))            # in the cyclotomic subgroup, order divides p^4 - p^2 + 1, not r
    wires = [bytes(576), ser.fp12_to_bytes(rnd), ser.fp12_to_bytes(cyc), ser.fp12_to_bytes(FP12_ONE), w[:576]]
    want = [ser.fp12_from_bytes(x)[0] for x in wires]
    assert want == [False, False, False, True, True]
    back, ok = eng.deserialize("gt", b"".join(wires))
    assert [bool(x) for x in ok] == want
    assert back[:3 * 576] == bytes(3 * 576) and back[4 * 576:] == g[:576]


def test_struct_serialisation_round_trip(eng):
    """CRS / Commit1 / Commit2 / EquProof in ark-serialize's compressed layout (u64-LE Vec lengths, two compressed
    points per commitment, 32 B LE scalars, 1-byte EquType): sizes, oracle-encoded bytes, round trips, rejection."""
    from groth_sahai_rs_b200 import api
    crs, _ = make_crs(1)
    key = api.CRS.from_bytes(crs_bytes(crs), eng)
    w = api.serialize_crs(key)
    assert len(w) == 1312
    want = ((2).to_bytes(8, "little") + b"".join(ser.g1_compress(p) for c in crs.u for p in c) +
            (2).to_bytes(8, "little") + b"".join(ser.g2_compress(p) for c in crs.v for p in c) +
            ser.g1_compress(crs.g1_gen) + ser.g2_compress(crs.g2_gen) + ser.fp12_to_bytes(crs.gt_gen))
    assert w == want
    assert api.deserialize_crs(w, eng).to_bytes() == crs_bytes(crs)
    rng = SeededRng(71)
    equ, xv, yv = random_instance(0, 3, 2, crs, rng)
    xr, yr, T = draw_rands(0, 3, 2, rng)
    pr = ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)
    c1 = api.Commit1([com1_b(c) for c in pr.xcoms.coms], [[fr_b(x) for x in r] for r in xr])
    w1 = api.serialize_commit(c1)
    assert len(w1) == 8 + 3 * 96 + 8 + 3 * (8 + 64)
    back = api.deserialize_commit(w1, 1, eng)
    assert back.coms == c1.coms and back.rand == c1.rand
    ep = pr.equ_proofs[0]
    p = api.EquProof([com2_b(c) for c in ep.pi], [com1_b(c) for c in ep.theta], 0, [[fr_b(x) for x in r] for r in T])
    wp = api.serialize_equ_proof(p, eng)
    assert len(wp) == 8 + 2 * 192 + 8 + 2 * 96 + 1 + 8 + 2 * (8 + 64)
    bp = api.deserialize_equ_proof(wp, eng)
    assert (bp.pi, bp.theta, bp.equ_type, bp.rand) == (p.pi, p.theta, 0, p.rand)
    bad = bytearray(wp)
    bad[8 + 2 * 192 + 8 + 2 * 96] = 4                                           # EquType byte out of range
    with pytest.raises(api.SerializationError):
        api.deserialize_equ_proof(bytes(bad), eng)
    bad = bytearray(wp)
    bad[8 + 5] ^= 0x55                                                          # corrupt an x-coordinate of pi
    with pytest.raises(api.SerializationError):
        api.deserialize_equ_proof(bytes(bad), eng)
    with pytest.raises(api.SerializationError):
        api.deserialize_equ_proof(wp[:-3], eng)


@pytest.mark.parametrize("ty", [0, 1, 2, 3])
def test_equation_serialisation(eng, ty):
    """PPE / MSMEG1 / MSMEG2 / QuadEqu in ark-serialize's compressed layout (statement.rs:117-185: a_consts, b_consts,
    gamma, target): bytes equal the oracle's element encodings, round trip (PartialEq), rejection, and the
    deserialised equation verifies the proof made for the original."""
    from groth_sahai_rs_b200 import api
    crs, _ = make_crs(1)
    key = api.CRS.from_bytes(crs_bytes(crs), eng)
    rng = SeededRng(300 + ty)
    m, n = 3, 2
    equ, xv, yv = random_instance(ty, m, n, crs, rng, zero_frac=0.3)
    asz, bsz = (96 if ty in (0, 1) else 32), (192 if ty in (0, 2) else 32)
    ab, bb = enc_A(ty, equ.a_consts), enc_B(ty, equ.b_consts)
    cls = [api.PPE, api.MSMEG1, api.MSMEG2, api.QuadEqu][ty]
    e = cls([ab[i:i + asz] for i in range(0, len(ab), asz)], [bb[i:i + bsz] for i in range(0, len(bb), bsz)],
            [[fr_b(g) for g in row] for row in equ.gamma], enc_T(ty, equ.target))
    w = api.serialize_equation(e, eng)
    enc_a = ser.g1_compress if ty in (0, 1) else ser.fr_to_bytes
    enc_b = ser.g2_compress if ty in (0, 2) else ser.fr_to_bytes
    enc_t = [ser.fp12_to_bytes, ser.g1_compress, ser.g2_compress, ser.fr_to_bytes][ty]
    u64 = lambda k: k.to_bytes(8, "little")
    want = (u64(n) + b"".join(enc_a(a) for a in equ.a_consts) + u64(m) + b"".join(enc_b(b) for b in equ.b_consts) +
            u64(m) + b"".join(u64(n) + b"".join(ser.fr_to_bytes(g) for g in row) for row in equ.gamma) +
            enc_t(equ.target))
    assert w == want
    back = api.deserialize_equation(w, ty, eng)
    assert type(back) is cls and back == e
    with pytest.raises(api.SerializationError):
        api.deserialize_equation(w[:-1], ty, eng)
    bad = bytearray(w)
    if ty == 0:
        bad[-48:] = P.to_bytes(48, "little")                                    # an Fp12 coefficient >= p
    elif ty == 3:
        bad[-1] = 0xFF                                                          # Fr >= r
    else:
        bad[-1] ^= 0x5A                                                         # x moved off the curve / out of the subgroup
    with pytest.raises(api.SerializationError):
        api.deserialize_equation(bytes(bad), ty, eng)
    xr, yr, T = draw_rands(ty, m, n, rng)
    pr = ogs.commit_and_prove(equ, xv, yv, crs, xr, yr, T)
    ep = pr.equ_proofs[0]
    cp = api.CProof(api.Commit1([com1_b(c) for c in pr.xcoms.coms], []), api.Commit2([com2_b(c) for c in pr.ycoms.coms], []),
                    [api.EquProof([com2_b(c) for c in ep.pi], [com1_b(c) for c in ep.theta], ty, [])])
    assert back.verify(cp, key) is True
    # an empty equation (no variables) keeps the u64 zero lengths
    e0 = cls([], [], [], enc_T(ty, equ.target))
    w0 = api.serialize_equation(e0, eng)
    assert w0[:24] == bytes(24) and api.deserialize_equation(w0, ty, eng) == e0


def test_uncompressed_encodings(eng):
    """serialize_uncompressed / deserialize_uncompressed (the reference round-trips both modes,
    data_structures.rs:1269-1309): bytes equal the oracle's, round trip, and the rejections of Validate::Yes."""
    rng = SeededRng(81)
    p1 = [G1_GEN, None] + [rng.g1() for _ in range(6)]
    w1 = eng.serialize("g1", b"".join(g1_b(p) for p in p1), compressed=False)
    assert w1 == b"".join(ser.g1_serialize_uncompressed(p) for p in p1)
    back, ok = eng.deserialize("g1", w1, compressed=False)
    assert ok == b"\x01" * len(p1) and back == b"".join(g1_b(p) for p in p1)
    p2 = [G2_GEN_FP2, None] + [rng.g2() for _ in range(4)]
    w2 = eng.serialize("g2", b"".join(g2_b(p) for p in p2), compressed=False)
    assert w2 == b"".join(ser.g2_serialize_uncompressed(p) for p in p2)
    back, ok = eng.deserialize("g2", w2, compressed=False)
    assert ok == b"\x01" * len(p2) and back == b"".join(g2_b(p) for p in p2)
    # rejections: compressed flag set, y off the curve, coordinate >= p, curve point outside the subgroup
    g = ser.g1_serialize_uncompressed(G1_GEN)
    bad = [bytes([g[0] | 0x80]) + g[1:], g[:95] + bytes([g[95] ^ 1]), P.to_bytes(48, "big") + g[48:]]
    x = 2
    while True:
        x += 1
        y = ser.fp_sqrt((x ** 3 + 4) % P)
        if y is not None and not ser.in_subgroup_g1((x, y)):
            break
    bad.append(x.to_bytes(48, "big") + y.to_bytes(48, "big"))
    back, ok = eng.deserialize("g1", b"".join(bad), compressed=False)
    assert ok == bytes(4) and back == bytes(96 * 4)
    for b_ in bad:
        assert ser.g1_deserialize_uncompressed(b_)[0] is False
    assert eng.deserialize("g1", bad[3], check_subgroup=False, compressed=False)[1] == b"\x01"
    g2w = ser.g2_serialize_uncompressed(G2_GEN_FP2)
    bad2 = [g2w[:191] + bytes([g2w[191] ^ 1]), g2w[:96] + P.to_bytes(48, "big") + g2w[144:]]
    assert eng.deserialize("g2", b"".join(bad2), compressed=False)[1] == bytes(2)
    assert all(ser.g2_deserialize_uncompressed(b_)[0] is False for b_ in bad2)
    # infinity must be canonical in the uncompressed form too: flag byte 0x40 followed by zeros only
    inf_cases = [bytes([0x40]) + bytes(95), bytes([0x60]) + bytes(95), bytes([0x40]) + bytes(94) + b"\x01"]
    back, ok = eng.deserialize("g1", b"".join(inf_cases), compressed=False)
    assert ok == b"\x01\x00\x00" and back == bytes(96 * 3)
    assert [ser.g1_deserialize_uncompressed(b_)[0] for b_ in inf_cases] == [True, False, False]
    inf2 = [bytes([0x40]) + bytes(191), bytes([0x40]) + bytes(100) + b"\x09" + bytes(90)]
    assert eng.deserialize("g2", b"".join(inf2), compressed=False)[1] == b"\x01\x00"
    assert [ser.g2_deserialize_uncompressed(b_)[0] for b_ in inf2] == [True, False]
