"""Synthetic Groth-Sahai workloads built ON THE GPU ENGINE (witnesses = multiples of the CRS generators, as
benches/bench.rs:314 does with crs.g1_gen.mul(Fr::rand)): shared by the GPU tests at the BASELINE.json shapes
and by tools/bench_configs.py.  `eng` is a groth_sahai_rs_b200.Engine with `eng._crs` = the oracle-form CRS."""
from gsutil import *  # noqa: F401,F403


def multiples_g1(eng, ks):
    g = g1_b(eng._crs.g1_gen)
    out = eng.com1_matmul(len(ks), 1, 1, b"".join(fr_b(k) for k in ks), g + g)
    return [out[i * 192:i * 192 + 96] for i in range(len(ks))]


def multiples_g2(eng, ks):
    g = g2_b(eng._crs.g2_gen)
    out = eng.com2_matmul(len(ks), 1, 1, b"".join(fr_b(k) for k in ks), g + g)
    return [out[i * 384:i * 384 + 192] for i in range(len(ks))]


def instance(eng, ty, m, n, rng):
    """A satisfied equation of type `ty` built on the GPU (witnesses = multiples of the generators)."""
    xs, ys = [rng.fr() for _ in range(m)], [rng.fr() for _ in range(n)]
    a, b = [rng.fr() for _ in range(n)], [rng.fr() for _ in range(m)]
    gam = [[rng.fr() for _ in range(n)] for _ in range(m)]
    val = (sum(a[j] * ys[j] for j in range(n)) + sum(xs[i] * b[i] for i in range(m)) +
           sum(gam[i][j] * xs[i] * ys[j] for i in range(m) for j in range(n))) % R
    g1A = ty in (0, 1)
    g2B = ty in (0, 2)
    X = b"".join(multiples_g1(eng, xs)) if g1A else frs_b(xs)
    A = b"".join(multiples_g1(eng, a)) if g1A else frs_b(a)
    Y = b"".join(multiples_g2(eng, ys)) if g2B else frs_b(ys)
    B = b"".join(multiples_g2(eng, b)) if g2B else frs_b(b)
    if ty == 0:
        T = eng.pairing(multiples_g1(eng, [val])[0], g2_b(eng._crs.g2_gen))
    elif ty == 1:
        T = multiples_g1(eng, [val])[0]
    elif ty == 2:
        T = multiples_g2(eng, [val])[0]
    else:
        T = fr_b(val)
    return A, B, frmat_b(gam), T, X, Y


def commit_prove(eng, ty, m, n, inst, rng):
    A, B, G, T, X, Y = inst
    cx = 2 if ty in (0, 1) else 1
    cy = 2 if ty in (0, 2) else 1
    xr = b"".join(fr_b(rng.fr()) for _ in range(m * cx))
    yr = b"".join(fr_b(rng.fr()) for _ in range(n * cy))
    Tr = b"".join(fr_b(rng.fr()) for _ in range(cx * cy))
    xc = eng.batch_commit_g1(X, xr) if ty in (0, 1) else eng.batch_commit_scalar_b1(X, xr)
    yc = eng.batch_commit_g2(Y, yr) if ty in (0, 2) else eng.batch_commit_scalar_b2(Y, yr)
    pi, th = eng.prove(ty, m, n, A, B, G, X, Y, xr, yr, Tr)
    return [A, B, G, T, xc, yc, pi, th]


def instance_many(eng, ty, m, n, E, rng):
    """E satisfied equations of type `ty` over ONE set of variables (a multi-equation statement, the C4 shape).
    Returns per-equation lists A, B, G, T and the shared X, Y, x_rand, y_rand plus per-equation T-randomness."""
    g1A, g2B = ty in (0, 1), ty in (0, 2)
    cx, cy = (2 if g1A else 1), (2 if g2B else 1)
    xs, ys = [rng.fr() for _ in range(m)], [rng.fr() for _ in range(n)]
    X = b"".join(multiples_g1(eng, xs)) if g1A else frs_b(xs)
    Y = b"".join(multiples_g2(eng, ys)) if g2B else frs_b(ys)
    a = [[rng.fr() for _ in range(n)] for _ in range(E)]
    b = [[rng.fr() for _ in range(m)] for _ in range(E)]
    gam = [[[rng.fr() for _ in range(n)] for _ in range(m)] for _ in range(E)]
    vals = [(sum(a[e][j] * ys[j] for j in range(n)) + sum(xs[i] * b[e][i] for i in range(m)) +
             sum(gam[e][i][j] * xs[i] * ys[j] for i in range(m) for j in range(n))) % R for e in range(E)]
    flat = lambda rows: [x for r in rows for x in r]
    Aall = multiples_g1(eng, flat(a)) if g1A else [fr_b(x) for x in flat(a)]
    Ball = multiples_g2(eng, flat(b)) if g2B else [fr_b(x) for x in flat(b)]
    if ty == 0:
        tg = eng.pairing(b"".join(multiples_g1(eng, vals)), g2_b(eng._crs.g2_gen) * E)
        T = [tg[576 * e:576 * (e + 1)] for e in range(E)]
    elif ty == 1:
        T = multiples_g1(eng, vals)
    elif ty == 2:
        T = multiples_g2(eng, vals)
    else:
        T = [fr_b(v) for v in vals]
    A = [b"".join(Aall[e * n:(e + 1) * n]) for e in range(E)]
    B = [b"".join(Ball[e * m:(e + 1) * m]) for e in range(E)]
    G = [frmat_b(gam[e]) for e in range(E)]
    xr = b"".join(fr_b(rng.fr()) for _ in range(m * cx))
    yr = b"".join(fr_b(rng.fr()) for _ in range(n * cy))
    Tr = [b"".join(fr_b(rng.fr()) for _ in range(cx * cy)) for _ in range(E)]
    return A, B, G, T, X, Y, xr, yr, Tr
