"""Canonical (de)serialisation oracle for the wire formats either side of the hot path (TEST INFRASTRUCTURE ONLY).

What it restates: arkworks' ``CanonicalSerialize / CanonicalDeserialize`` as derived by the reference for
``CRS`` (generator.rs:35), ``Commit1/Commit2`` (prover/commit.rs:18-28), ``EquProof`` (prover/prove.rs:55-61)
and the equations (statement.rs:117-185).  The arithmetic-bearing part is the point encoding; ark-bls12-381 >= 0.4
(``curves/util.rs``, not vendored here -- SURVEY.md §8c) uses the zcash / IETF BLS12-381 format:

* G1: 48 B compressed = big-endian x, 96 B uncompressed = x || y; G2: 96 B compressed = x.c1 || x.c0 (big-endian),
  192 B uncompressed = x.c1 || x.c0 || y.c1 || y.c0.
* top three bits of byte 0: 0x80 compressed, 0x40 infinity, 0x20 "y is the lexicographically largest of {y, -y}"
  (Fp2 ordered with c1 most significant); infinity = flag byte followed by zeros.
* deserialisation (Validate::Yes) rejects x >= p, x with no point, points outside the order-r subgroup; a set
  infinity flag yields the identity ONLY in its canonical form (sort flag clear, every other bit zero: ark-bls12-381's
  EncodingFlags::get_flags rejects sort + infinity and read_g*_compressed / _uncompressed require zero coordinates).
* GT (PairingOutput): besides canonical coefficients, Valid::check requires f^r == 1 (ark-ec 0.5 pairing.rs).
* Fr: 32 B little-endian canonical integer (< r); Fp12/GT: 12 x 48 B little-endian canonical, tower order.

PARITY STATUS: **unpinned against arkworks bits** (no Rust toolchain).  Pinned against the public compressed
generator encodings of the zcash / IETF specification (tests/test_serialize.py): G1 ``97f1d3a7...c6bb`` and G2
``93e02b60...bdb8``, which decode to the generator coordinates every BLS12-381 library publishes.
"""
from .bls12_381 import P, R, Fp2, G1, G2

HALF_P = (P - 1) // 2


# ---------------------------------------------------------------- field helpers
def fp_sqrt(a: int):
    """Square root in Fp (p = 3 mod 4) or None."""
    a %= P
    y = pow(a, (P + 1) // 4, P)
    return y if y * y % P == a else None


def fp2_sqrt(a: Fp2):
    """Square root in Fp2 = Fp[u]/(u^2+1) by the norm method, or None."""
    a0, a1 = a.c0 % P, a.c1 % P
    if a1 == 0:
        s = fp_sqrt(a0)
        if s is not None:
            return Fp2(s, 0)
        s = fp_sqrt(-a0 % P)          # a0 a non-residue: sqrt = u * sqrt(-a0)
        return None if s is None else Fp2(0, s)
    n = fp_sqrt((a0 * a0 + a1 * a1) % P)
    if n is None:
        return None
    inv2 = pow(2, P - 2, P)
    for nn in (n, -n % P):
        x0 = fp_sqrt((a0 + nn) * inv2 % P)
        if x0 is not None and x0 != 0:
            x1 = a1 * pow(2 * x0, P - 2, P) % P
            r = Fp2(x0, x1)
            if r * r == Fp2(a0, a1):
                return r
    return None


def fp2_is_largest(y: Fp2) -> bool:
    """y > -y with Fp2 ordered (c1, c0) lexicographically, as ark-ff's Ord for Fp2 / the zcash spec."""
    c0, c1 = y.c0 % P, y.c1 % P
    if c1 != 0:
        return c1 > HALF_P
    return c0 > HALF_P


def in_subgroup_g1(pt) -> bool:
    return pt is None or G1.mul(pt, R) is None


def in_subgroup_g2(pt) -> bool:
    return pt is None or G2.mul(pt, R) is None


# the endomorphism membership tests ark-bls12-381 runs (Scott, ePrint 2021/1130 Sections 6 and 4); the constants
# are fixed by phi(G) = -[x^2] G and psi(G2) = [x] G2 on the generators
from .bls12_381 import X_ABS, G1_GEN, G2_GEN_FP2  # noqa: E402
_XI = Fp2(1, 1)
PSI_CX = _XI.pow((P - 1) // 3).inv()
PSI_CY = _XI.pow((P - 1) // 2).inv()
ENDO_BETA = 0x5F19672FDF76CE51BA69C6076A0F77EADDB3A93BE6F89688DE17D813620A00022E01FFFFFFFEFFFE


def in_subgroup_g1_fast(pt) -> bool:
    if pt is None:
        return True
    t1 = G1.mul(pt, X_ABS)
    if t1 == pt:
        return False
    t2 = G1.mul(t1, X_ABS)
    return t2 == (ENDO_BETA * pt[0] % P, -pt[1] % P)


def in_subgroup_g2_fast(pt) -> bool:
    if pt is None:
        return True
    t = G2.mul(pt, X_ABS)
    psi = (pt[0].conj() * PSI_CX, pt[1].conj() * PSI_CY)
    return t == G2.neg(psi)


# ---------------------------------------------------------------- G1
def g1_compress(pt) -> bytes:
    if pt is None:
        return bytes([0xC0]) + bytes(47)
    x, y = pt[0] % P, pt[1] % P
    b = bytearray(x.to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if y > HALF_P else 0)
    return bytes(b)


def g1_decompress(b: bytes):
    """-> (ok, point).  ok False: not compressed, x >= p, no such point, or not in the order-r subgroup."""
    assert len(b) == 48
    flags = b[0]
    if not flags & 0x80:
        return False, None
    if flags & 0x40:           # infinity must be canonical: sort flag clear, x bytes zero (EncodingFlags::get_flags)
        return (True, None) if (flags & 0x3F) == 0 and not any(b[1:]) else (False, None)
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= P:
        return False, None
    y = fp_sqrt((x * x * x + 4) % P)
    if y is None:
        return False, None
    if (y > HALF_P) != bool(flags & 0x20):
        y = -y % P
    pt = (x, y)
    return (True, pt) if in_subgroup_g1(pt) else (False, None)


def g1_serialize_uncompressed(pt) -> bytes:
    if pt is None:
        return bytes([0x40]) + bytes(95)
    return (pt[0] % P).to_bytes(48, "big") + (pt[1] % P).to_bytes(48, "big")


def g1_deserialize_uncompressed(b: bytes):
    assert len(b) == 96
    if b[0] & 0x80:
        return False, None
    if b[0] & 0x40:
        return (True, None) if (b[0] & 0x3F) == 0 and not any(b[1:]) else (False, None)
    if b[0] & 0x20:
        return False, None
    x, y = int.from_bytes(b[:48], "big"), int.from_bytes(b[48:], "big")
    if x >= P or y >= P or (y * y - x * x * x - 4) % P != 0:
        return False, None
    return (True, (x, y)) if in_subgroup_g1((x, y)) else (False, None)


# ---------------------------------------------------------------- G2
def g2_serialize_uncompressed(pt) -> bytes:
    if pt is None:
        return bytes([0x40]) + bytes(191)
    x, y = pt
    return b"".join((v % P).to_bytes(48, "big") for v in (x.c1, x.c0, y.c1, y.c0))


def g2_deserialize_uncompressed(b: bytes):
    assert len(b) == 192
    if b[0] & 0x80:
        return False, None
    if b[0] & 0x40:
        return (True, None) if (b[0] & 0x3F) == 0 and not any(b[1:]) else (False, None)
    if b[0] & 0x20:
        return False, None
    v = [int.from_bytes(b[48 * i:48 * (i + 1)], "big") for i in range(4)]
    if any(t >= P for t in v):
        return False, None
    x, y = Fp2(v[1], v[0]), Fp2(v[3], v[2])
    if not y * y == x * x * x + Fp2(4, 4):
        return False, None
    return (True, (x, y)) if in_subgroup_g2((x, y)) else (False, None)


def g2_compress(pt) -> bytes:
    if pt is None:
        return bytes([0xC0]) + bytes(95)
    x, y = pt
    b = bytearray((x.c1 % P).to_bytes(48, "big") + (x.c0 % P).to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if fp2_is_largest(y) else 0)
    return bytes(b)


def g2_decompress(b: bytes):
    assert len(b) == 96
    flags = b[0]
    if not flags & 0x80:
        return False, None
    if flags & 0x40:
        return (True, None) if (flags & 0x3F) == 0 and not any(b[1:]) else (False, None)
    c1 = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:48], "big")
    c0 = int.from_bytes(b[48:96], "big")
    if c1 >= P or c0 >= P:
        return False, None
    x = Fp2(c0, c1)
    y = fp2_sqrt(x * x * x + Fp2(4, 4))
    if y is None:
        return False, None
    if fp2_is_largest(y) != bool(flags & 0x20):
        y = Fp2(0, 0) - y
    pt = (x, y)
    return (True, pt) if in_subgroup_g2(pt) else (False, None)


# ---------------------------------------------------------------- scalars / GT
def fr_to_bytes(a: int) -> bytes:
    return (a % R).to_bytes(32, "little")


def fr_from_bytes(b: bytes):
    v = int.from_bytes(b, "little")
    return (v < R), (v if v < R else None)


def fp12_to_bytes(f) -> bytes:
    out = b""
    for c6 in (f.c0, f.c1):
        for c2 in (c6.c0, c6.c1, c6.c2):
            out += (c2.c0 % P).to_bytes(48, "little") + (c2.c1 % P).to_bytes(48, "little")
    return out


def fp12_from_bytes(b: bytes):
    """-> (ok, Fp12).  CanonicalDeserialize of a PairingOutput with Validate::Yes: 12 canonical Fp coefficients in tower
    order, then Valid::check: f^r == 1 (zero and elements outside the order-r subgroup are rejected)."""
    from .bls12_381 import Fp6, Fp12, FP12_ONE
    assert len(b) == 576
    v = [int.from_bytes(b[48 * i:48 * (i + 1)], "little") for i in range(12)]
    if any(t >= P for t in v):
        return False, None
    c = [Fp2(v[2 * i], v[2 * i + 1]) for i in range(6)]
    f = Fp12(Fp6(c[0], c[1], c[2]), Fp6(c[3], c[4], c[5]))
    return (True, f) if f.pow(R) == FP12_ONE else (False, None)
