"""Conversions between oracle objects (Python ints / Fp2 / tuples) and the C-ABI byte layout
(arkworks in-memory Montgomery limbs, little-endian; affine identity = all-zero bytes)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bls12_381 import (Fp2, Fp6, Fp12, fp_to_mont_bytes, fp_from_mont_bytes,
                              fr_to_mont_bytes, fr_from_mont_bytes)

def fp_b(a): return fp_to_mont_bytes(a)
def fp_i(b): return fp_from_mont_bytes(b)
def fr_b(a): return fr_to_mont_bytes(a)
def fr_i(b): return fr_from_mont_bytes(b)
def fp2_b(a): return fp_b(a.c0) + fp_b(a.c1)
def fp2_i(b): return Fp2(fp_i(b[:48]), fp_i(b[48:96]))
def fp6_b(a): return fp2_b(a.c0) + fp2_b(a.c1) + fp2_b(a.c2)
def fp6_i(b): return Fp6(fp2_i(b[0:96]), fp2_i(b[96:192]), fp2_i(b[192:288]))
def fp12_b(a): return fp6_b(a.c0) + fp6_b(a.c1)
def fp12_i(b): return Fp12(fp6_i(b[:288]), fp6_i(b[288:576]))
def g1_b(p): return bytes(96) if p is None else fp_b(p[0]) + fp_b(p[1])
def g1_i(b): return None if b == bytes(96) else (fp_i(b[:48]), fp_i(b[48:96]))
def g2_b(p): return bytes(192) if p is None else fp2_b(p[0]) + fp2_b(p[1])
def g2_i(b): return None if b == bytes(192) else (fp2_i(b[:96]), fp2_i(b[96:192]))
def com1_b(c): return g1_b(c[0]) + g1_b(c[1])
def com1_i(b): return (g1_i(b[:96]), g1_i(b[96:192]))
def com2_b(c): return g2_b(c[0]) + g2_b(c[1])
def com2_i(b): return (g2_i(b[:192]), g2_i(b[192:384]))
def comt_b(c): return b"".join(fp12_b(x) for x in c)
def comt_i(b): return [fp12_i(b[576 * i:576 * (i + 1)]) for i in range(4)]
def frs_b(xs): return b"".join(fr_b(x) for x in xs)
def frmat_b(m): return b"".join(fr_b(x) for row in m for x in row)
