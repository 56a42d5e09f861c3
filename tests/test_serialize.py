"""CPU tests of the serialisation oracle (oracle/serialize.py): pinned by the public compressed encodings of
the BLS12-381 generators (zcash / IETF), round trips, flag handling and rejection of invalid encodings."""
import random

from oracle import serialize as ser
from oracle.bls12_381 import P, R, G1, G2, G1_GEN, G2_GEN_FP2, Fp2, g1_mul, g2_mul

# the generator encodings every BLS12-381 library publishes (zcash pairing crate docs, IETF BLS draft appendix)
G1_GEN_COMPRESSED = bytes.fromhex(
    "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb")
G2_GEN_COMPRESSED = bytes.fromhex(
    "93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
    "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")


def test_generator_vectors():
    assert ser.g1_compress(G1_GEN) == G1_GEN_COMPRESSED
    assert ser.g2_compress(G2_GEN_FP2) == G2_GEN_COMPRESSED
    assert ser.g1_decompress(G1_GEN_COMPRESSED) == (True, G1_GEN)
    ok, q = ser.g2_decompress(G2_GEN_COMPRESSED)
    assert ok and q == G2_GEN_FP2
    # the negated generators flip exactly the sort bit
    neg = ser.g1_compress(G1.neg(G1_GEN))
    assert neg[0] == G1_GEN_COMPRESSED[0] | 0x20 and neg[1:] == G1_GEN_COMPRESSED[1:]
    neg2 = ser.g2_compress(G2.neg(G2_GEN_FP2))
    assert neg2[0] == G2_GEN_COMPRESSED[0] | 0x20 and neg2[1:] == G2_GEN_COMPRESSED[1:]


def test_round_trips_and_infinity():
    rnd = random.Random(7)
    for _ in range(6):
        p = g1_mul(G1_GEN, rnd.randrange(R))
        assert ser.g1_decompress(ser.g1_compress(p)) == (True, p)
    for _ in range(3):
        q = g2_mul(G2_GEN_FP2, rnd.randrange(R))
        ok, back = ser.g2_decompress(ser.g2_compress(q))
        assert ok and back == q
    assert ser.g1_compress(None) == bytes([0xC0]) + bytes(47)
    assert ser.g1_decompress(bytes([0xC0]) + bytes(47)) == (True, None)
    assert ser.g2_decompress(bytes([0xC0]) + bytes(95)) == (True, None)
    # non-canonical infinity encodings are rejected (ark-bls12-381: sort flag + infinity, or non-zero coordinate bytes)
    assert ser.g1_decompress(bytes([0xE0]) + bytes(47)) == (False, None)
    assert ser.g1_decompress(bytes([0xC0]) + bytes(46) + b"\x01") == (False, None)
    assert ser.g2_decompress(bytes([0xE0]) + bytes(95)) == (False, None)
    assert ser.g2_decompress(bytes([0xC0]) + bytes(60) + b"\x01" + bytes(34)) == (False, None)
    assert ser.g1_deserialize_uncompressed(bytes([0x40]) + bytes(95)) == (True, None)
    assert ser.g1_deserialize_uncompressed(bytes([0x40]) + bytes(94) + b"\x02") == (False, None)
    assert ser.g2_deserialize_uncompressed(bytes([0x60]) + bytes(191)) == (False, None)


def test_gt_valid_check():
    """PairingOutput's Valid::check on deserialisation: f^r == 1."""
    from oracle.bls12_381 import Fp6, Fp12, FP12_ONE, pairing
    e = pairing(G1_GEN, G2_GEN_FP2)
    assert ser.fp12_from_bytes(ser.fp12_to_bytes(e)) == (True, e)
    assert ser.fp12_from_bytes(ser.fp12_to_bytes(FP12_ONE))[0] is True
    assert ser.fp12_from_bytes(bytes(576))[0] is False
    two = Fp12(Fp6(Fp2(2, 0), Fp2(0, 0), Fp2(0, 0)), Fp6(Fp2(0, 0), Fp2(0, 0), Fp2(0, 0)))
    assert ser.fp12_from_bytes(ser.fp12_to_bytes(two))[0] is False          # 2 has order dividing p - 1, not r
    bad = bytearray(ser.fp12_to_bytes(e))
    bad[48:96] = P.to_bytes(48, "little")
    assert ser.fp12_from_bytes(bytes(bad))[0] is False


def test_rejections():
    # not flagged compressed
    assert ser.g1_decompress(bytes([G1_GEN_COMPRESSED[0] & 0x7F]) + G1_GEN_COMPRESSED[1:])[0] is False
    # x >= p
    bad = bytearray(P.to_bytes(48, "big"))
    bad[0] |= 0x80
    assert ser.g1_decompress(bytes(bad))[0] is False
    # an x with no point on the curve, and one on the curve but outside the order-r subgroup
    x = 1
    seen_nopoint = seen_offgroup = False
    while not (seen_nopoint and seen_offgroup):
        x += 1
        enc = bytearray(x.to_bytes(48, "big"))
        enc[0] |= 0x80
        ok, _ = ser.g1_decompress(bytes(enc))
        y = ser.fp_sqrt((x ** 3 + 4) % P)
        if y is None:
            assert not ok
            seen_nopoint = True
        elif not ser.in_subgroup_g1((x, y)):
            assert not ok
            seen_offgroup = True
    # G2: a curve point outside the subgroup
    c0 = 0
    while True:
        c0 += 1
        xx = Fp2(c0, 1)
        yy = ser.fp2_sqrt(xx * xx * xx + Fp2(4, 4))
        if yy is not None and not ser.in_subgroup_g2((xx, yy)):
            enc = bytearray((1).to_bytes(48, "big") + c0.to_bytes(48, "big"))
            enc[0] |= 0x80
            assert ser.g2_decompress(bytes(enc))[0] is False
            break


def test_fp2_sqrt_and_scalars():
    rnd = random.Random(9)
    for _ in range(20):
        a = Fp2(rnd.randrange(P), rnd.randrange(P))
        s = ser.fp2_sqrt(a * a)
        assert s is not None and s * s == a * a
    for a in (Fp2(5, 0), Fp2(0, 7), Fp2(P - 1, 0)):           # c1 = 0 branches (residue / non-residue real part)
        s = ser.fp2_sqrt(a)
        assert s is None or s * s == a
    assert ser.fp2_sqrt(Fp2(P - 1, 0)) is not None              # -1 = u^2
    assert ser.fr_from_bytes(ser.fr_to_bytes(12345)) == (True, 12345)
    assert ser.fr_from_bytes(R.to_bytes(32, "little"))[0] is False


def test_endomorphism_membership_tests_agree_with_definition():
    """The 64-bit endomorphism tests (what arkworks and serial.cu run) == [r]P = O, inside and outside G1 / G2."""
    rnd = random.Random(11)
    for _ in range(3):
        assert ser.in_subgroup_g1_fast(g1_mul(G1_GEN, rnd.randrange(1, R)))
        assert ser.in_subgroup_g2_fast(g2_mul(G2_GEN_FP2, rnd.randrange(1, R)))
    found = 0
    x = 1
    while found < 6:
        x += 1
        y = ser.fp_sqrt((x ** 3 + 4) % P)
        if y is None:
            continue
        slow = ser.in_subgroup_g1((x, y))
        assert ser.in_subgroup_g1_fast((x, y)) == slow
        found += 0 if slow else 1
    # a point of the subgroup plus a point of small cofactor order is outside: h1 = (x-1)^2/3 has the factor 3
    found = 0
    c0 = 0
    while found < 4:
        c0 += 1
        xx = Fp2(c0, 2)
        yy = ser.fp2_sqrt(xx * xx * xx + Fp2(4, 4))
        if yy is None:
            continue
        slow = ser.in_subgroup_g2((xx, yy))
        assert ser.in_subgroup_g2_fast((xx, yy)) == slow
        found += 0 if slow else 1
