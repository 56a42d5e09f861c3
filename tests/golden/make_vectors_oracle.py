"""Writes tests/golden/vectors_oracle.jsonl: the record format of tests/vectors.py filled in by the big-int oracle
(source = "oracle": these do NOT pin anything to arkworks; tools/gen_vectors.rs writes the same records from the
reference).  Run from the repo root:  python tests/golden/make_vectors_oracle.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

from gsutil import SeededRng, make_crs, random_instance, draw_rands  # noqa: E402
from oracle import serialize as ser  # noqa: E402
import vectors as V  # noqa: E402

frh = lambda x: ser.fr_to_bytes(x).hex()


def main():
    recs = []
    crs, (p1, p2, a1, a2, t1, t2) = make_crs(11)
    recs.append({"kind": "crs", "draws": {"p1": ser.g1_compress(p1).hex(), "p2": ser.g2_compress(p2).hex(),
                                           "fr": [frh(x) for x in (a1, a2, t1, t2)]}})
    crs_h = V.crs_e(crs).hex()
    rng = SeededRng(12)
    m = 2
    for kind, gen, enc, cols in (("commit_g1", rng.g1, ser.g1_compress, 2), ("commit_g2", rng.g2, ser.g2_compress, 2),
                                 ("commit_b1", rng.fr, ser.fr_to_bytes, 1), ("commit_b2", rng.fr, ser.fr_to_bytes, 1)):
        vs = [gen() for _ in range(m)]
        if kind in ("commit_g1", "commit_g2"):
            vs[1] = None                                  # the identity as a committed value
        rand = [[rng.fr() for _ in range(cols)] for _ in range(m)]
        recs.append({"kind": kind, "crs": crs_h, "vars": [enc(v).hex() for v in vs], "rand": V.frm_h(rand)})
    for ty in range(4):
        r = SeededRng(20 + ty)
        equ, xv, yv = random_instance(ty, 2, 1, crs, r, zero_frac=0.3)
        xr, yr, T = draw_rands(ty, 2, 1, r)
        recs.append({"kind": "prove", "equ_type": ty, "crs": crs_h, "equation": V.equation_e(equ).hex(),
                     "xvars": [V.A_ENC[ty](v).hex() for v in xv], "yvars": [V.B_ENC[ty](v).hex() for v in yv],
                     "xrand": V.frm_h(xr), "yrand": V.frm_h(yr), "T": V.frm_h(T)})
    xs = [(rng.g1(), rng.g1()), (None, rng.g1())]
    ys = [(rng.g2(), rng.g2()), (rng.g2(), None)]
    recs.append({"kind": "pairing_sum", "xs": [V.com1_e(x).hex() for x in xs], "ys": [V.com2_e(y).hex() for y in ys]})
    with open(V.path("oracle"), "w") as f:
        for rec in recs:
            rec["source"] = "oracle"
            rec.update(V.oracle_outputs(rec))
            f.write(json.dumps(rec) + "\n")
            print(rec["kind"], rec.get("equ_type", ""), "ok", flush=True)


if __name__ == "__main__":
    main()
