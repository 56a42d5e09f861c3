"""Cross-implementation vector files (tests/golden/vectors_*.jsonl): one JSON record per reference API call with EVERY
input, every host RNG draw (in the reference's draw order) and every output in ark-serialize's compressed layout, hex.

Two producers write the same format:
  * tools/gen_vectors.rs      -- the reference itself (groth-sahai-rs + ark-bls12-381); needs cargo, which this image lacks.
                                 Its output, dropped at tests/golden/vectors_arkworks.jsonl, PINS the oracle and the CUDA path
                                 to arkworks' bits ("parity unpinned" until then).
  * tests/golden/make_vectors_oracle.py -- the big-int oracle; committed as tests/golden/vectors_oracle.jsonl so the
                                 replay code below is exercised today.
Two consumers replay a record: `oracle_outputs` (CPU) and `api_outputs` (CUDA path through api.py / the C ABI).

Record kinds and fields (all values hex of serialize_compressed bytes; Fr = 32 B LE, Matrix<Fr> as lists of rows):
  crs            draws {p1, p2, fr[4]=a1,a2,t1,t2}                                   -> crs
  commit_g1/g2   crs, vars[], rand[m][2]      (commit.rs:78-100, 178-200)            -> commit   (Commit1/2 struct)
  commit_b1/b2   crs, vars[] (Fr), rand[m][1] (commit.rs:125-156, 225-256)           -> commit
  prove          equ_type, crs, equation, xvars[], yvars[], xrand, yrand, T          -> xcoms, ycoms, proof, verify
  pairing_sum    xs[] (Com1 = 2 points), ys[] (Com2)   (data_structures.rs:494-502)  -> comt[4]
"""
import json
import os

from oracle import gs as ogs
from oracle import serialize as ser
from oracle.bls12_381 import Fp2, Fp6, Fp12

HERE = os.path.dirname(os.path.abspath(__file__))
H = bytes.fromhex


def path(source): return os.path.join(HERE, "golden", f"vectors_{source}.jsonl")


def load(source):
    p = path(source)
    return [json.loads(l) for l in open(p) if l.strip()] if os.path.exists(p) else None


def u64(n): return int(n).to_bytes(8, "little")


# ------------------------------------------------------------------ oracle-side struct encoders / decoders
def _dec(fn, b, what):
    ok, v = fn(b)
    assert ok, f"invalid {what} in vector file"
    return v


def g1_d(h): return _dec(ser.g1_decompress, H(h), "G1")
def g2_d(h): return _dec(ser.g2_decompress, H(h), "G2")
def fr_d(h): return _dec(ser.fr_from_bytes, H(h), "Fr")


def gt_d(h):
    b = H(h)
    c = [int.from_bytes(b[48 * i:48 * i + 48], "little") for i in range(12)]
    f2 = [Fp2(c[2 * i], c[2 * i + 1]) for i in range(6)]
    return Fp12(Fp6(f2[0], f2[1], f2[2]), Fp6(f2[3], f2[4], f2[5]))


def com1_e(c): return ser.g1_compress(c[0]) + ser.g1_compress(c[1])
def com2_e(c): return ser.g2_compress(c[0]) + ser.g2_compress(c[1])
def com1_d(h): return (g1_d(h[:96]), g1_d(h[96:]))
def com2_d(h): return (g2_d(h[:192]), g2_d(h[192:]))
def vec_e(items, f): return u64(len(items)) + b"".join(f(x) for x in items)
def mat_e(rows): return u64(len(rows)) + b"".join(vec_e(r, ser.fr_to_bytes) for r in rows)


def crs_e(crs):
    return (vec_e(crs.u, com1_e) + vec_e(crs.v, com2_e) + ser.g1_compress(crs.g1_gen) + ser.g2_compress(crs.g2_gen) +
            ser.fp12_to_bytes(crs.gt_gen))


def crs_d(h):
    b = H(h)
    assert len(b) == 1312 and b[:8] == u64(2) and b[200:208] == u64(2)
    u = [com1_d(b[8:104].hex()), com1_d(b[104:200].hex())]
    v = [com2_d(b[208:400].hex()), com2_d(b[400:592].hex())]
    return ogs.CRS(u=u, v=v, g1_gen=g1_d(b[592:640].hex()), g2_gen=g2_d(b[640:736].hex()), gt_gen=gt_d(b[736:].hex()))


def commit_e(c, which): return vec_e(c.coms, com1_e if which == 1 else com2_e) + mat_e(c.rand)


def proof_e(ep): return vec_e(ep.pi, com2_e) + vec_e(ep.theta, com1_e) + bytes([ep.equ_type]) + mat_e(ep.rand)


A_ENC = {0: ser.g1_compress, 1: ser.g1_compress, 2: ser.fr_to_bytes, 3: ser.fr_to_bytes}
B_ENC = {0: ser.g2_compress, 1: ser.fr_to_bytes, 2: ser.g2_compress, 3: ser.fr_to_bytes}
T_ENC = {0: ser.fp12_to_bytes, 1: ser.g1_compress, 2: ser.g2_compress, 3: ser.fr_to_bytes}
A_DEC = {0: g1_d, 1: g1_d, 2: fr_d, 3: fr_d}
B_DEC = {0: g2_d, 1: fr_d, 2: g2_d, 3: fr_d}
T_DEC = {0: gt_d, 1: g1_d, 2: g2_d, 3: fr_d}
A_SZ = {0: 48, 1: 48, 2: 32, 3: 32}
B_SZ = {0: 96, 1: 32, 2: 96, 3: 32}


def equation_e(e):
    ty = e.equ_type
    return vec_e(e.a_consts, A_ENC[ty]) + vec_e(e.b_consts, B_ENC[ty]) + mat_e(e.gamma) + T_ENC[ty](e.target)


def equation_d(h, ty):
    b, o = H(h), 0

    def take(n):
        nonlocal o
        o += n
        assert o <= len(b)
        return b[o - n:o]

    def vec(sz, dec): return [dec(take(sz).hex()) for _ in range(int.from_bytes(take(8), "little"))]
    a = vec(A_SZ[ty], A_DEC[ty])
    bc = vec(B_SZ[ty], B_DEC[ty])
    gamma = [vec(32, fr_d) for _ in range(int.from_bytes(take(8), "little"))]
    t = T_DEC[ty](b[o:].hex())
    return ogs.Equation(ty, a, bc, gamma, t)


def frm(rows): return [[fr_d(x) for x in r] for r in rows]
def frm_h(rows): return [[ser.fr_to_bytes(x).hex() for x in r] for r in rows]


# ------------------------------------------------------------------ replay through the oracle (CPU)
def oracle_outputs(rec):
    k = rec["kind"]
    if k == "crs":
        d = rec["draws"]
        return {"crs": crs_e(ogs.generate_crs(g1_d(d["p1"]), g2_d(d["p2"]), *[fr_d(x) for x in d["fr"]])).hex()}
    if k.startswith("commit_"):
        crs, rand = crs_d(rec["crs"]), frm(rec["rand"])
        fn, dec, which = {"commit_g1": (ogs.batch_commit_g1, g1_d, 1), "commit_g2": (ogs.batch_commit_g2, g2_d, 2),
                          "commit_b1": (ogs.batch_commit_scalar_to_b1, fr_d, 1),
                          "commit_b2": (ogs.batch_commit_scalar_to_b2, fr_d, 2)}[k]
        return {"commit": commit_e(fn([dec(v) for v in rec["vars"]], crs, rand), which).hex()}
    if k == "prove":
        ty, crs = rec["equ_type"], crs_d(rec["crs"])
        equ = equation_d(rec["equation"], ty)
        xv = [A_DEC[ty](v) for v in rec["xvars"]]
        yv = [B_DEC[ty](v) for v in rec["yvars"]]
        cp = ogs.commit_and_prove(equ, xv, yv, crs, frm(rec["xrand"]), frm(rec["yrand"]), frm(rec["T"]))
        return {"xcoms": commit_e(cp.xcoms, 1).hex(), "ycoms": commit_e(cp.ycoms, 2).hex(),
                "proof": proof_e(cp.equ_proofs[0]).hex(), "verify": ogs.verify(equ, cp, crs)}
    if k == "pairing_sum":
        out = ogs.comt_pairing_sum([com1_d(x) for x in rec["xs"]], [com2_d(y) for y in rec["ys"]])
        return {"comt": [ser.fp12_to_bytes(g).hex() for g in out]}
    raise ValueError(k)


# ------------------------------------------------------------------ replay through api.py / the C ABI (GPU)
class ListRng:
    """Hands out recorded draws in ABI form; the call order must be the reference's or the lists run dry."""

    def __init__(self, fr=(), g1=(), g2=()):
        self._fr, self._g1, self._g2 = list(fr), list(g1), list(g2)

    def fr(self): return self._fr.pop(0)
    def g1(self): return self._g1.pop(0)
    def g2(self): return self._g2.pop(0)
    def dry(self): return not (self._fr or self._g1 or self._g2)


def api_outputs(rec, api, eng):
    def de(kind, hexes):
        if not hexes:
            return []
        out, ok = eng.deserialize(kind, b"".join(H(x) for x in hexes))
        assert ok == b"\x01" * len(hexes)
        size = len(out) // len(hexes)
        return [out[i * size:(i + 1) * size] for i in range(len(hexes))]

    def flat_fr(*mats): return de("fr", [x for m in mats for r in m for x in r])
    k = rec["kind"]
    if k == "crs":
        d = rec["draws"]
        rng = ListRng(de("fr", d["fr"]), de("g1", [d["p1"]]), de("g2", [d["p2"]]))
        crs = api.CRS.generate_crs(rng, eng)
        assert rng.dry()
        return {"crs": api.serialize_crs(crs).hex()}
    if k.startswith("commit_"):
        key = api.deserialize_crs(H(rec["crs"]), eng)
        fn, kind = {"commit_g1": (api.batch_commit_G1, "g1"), "commit_g2": (api.batch_commit_G2, "g2"),
                    "commit_b1": (api.batch_commit_scalar_to_B1, "fr"), "commit_b2": (api.batch_commit_scalar_to_B2, "fr")}[k]
        rng = ListRng(flat_fr(rec["rand"]))
        c = fn(de(kind, rec["vars"]), key, rng)
        assert rng.dry()
        return {"commit": api.serialize_commit(c, eng).hex()}
    if k == "prove":
        ty = rec["equ_type"]
        key = api.deserialize_crs(H(rec["crs"]), eng)
        equ = api.deserialize_equation(H(rec["equation"]), ty, eng)
        xv = de("g1" if ty in (0, 1) else "fr", rec["xvars"])
        yv = de("g2" if ty in (0, 2) else "fr", rec["yvars"])
        rng = ListRng(flat_fr(rec["xrand"], rec["yrand"], rec["T"]))      # draw order x -> y -> T (prove.rs:82-88)
        cp = equ.commit_and_prove(xv, yv, key, rng)
        assert rng.dry()
        return {"xcoms": api.serialize_commit(cp.xcoms, eng).hex(), "ycoms": api.serialize_commit(cp.ycoms, eng).hex(),
                "proof": api.serialize_equ_proof(cp.equ_proofs[0], eng).hex(), "verify": equ.verify(cp, key)}
    if k == "pairing_sum":
        xs = [a + b for a, b in zip(*[iter(de("g1", [x[i:i + 96] for x in rec["xs"] for i in (0, 96)]))] * 2)]
        ys = [a + b for a, b in zip(*[iter(de("g2", [y[i:i + 192] for y in rec["ys"] for i in (0, 192)]))] * 2)]
        t = api.ComT.pairing_sum(xs, ys, eng)
        w = eng.serialize("gt", t)
        return {"comt": [w[576 * i:576 * (i + 1)].hex() for i in range(4)]}
    raise ValueError(k)


OUTPUT_KEYS = {"crs": ["crs"], "commit_g1": ["commit"], "commit_g2": ["commit"], "commit_b1": ["commit"],
               "commit_b2": ["commit"], "prove": ["xcoms", "ycoms", "proof", "verify"], "pairing_sum": ["comt"]}
