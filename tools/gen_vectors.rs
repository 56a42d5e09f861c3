//! Dumps outputs of the REFERENCE (jdwhite48/groth-sahai-rs on ark-bls12-381) in the record format of
//! tests/vectors.py, one JSON object per line.  This image has no Rust toolchain, so the file is shipped
//! unbuilt; anyone with cargo closes the "parity unpinned" gap of DESIGN.md §2 with:
//!
//!     cp tools/gen_vectors.rs <groth-sahai-rs checkout>/examples/gen_vectors.rs
//!     (cd <checkout> && cargo run --release --example gen_vectors) > tests/golden/vectors_arkworks.jsonl
//!     python -m pytest tests/test_vectors.py            # CPU: pins the oracle;  -m gpu: pins the CUDA path
//!
//! (ark-bls12-381 is already a dev-dependency of the reference, which is all an example needs.)
//! Every value is `serialize_compressed` bytes in hex.  Private randomness (`Commit*.rand`, `EquProof.rand`) is read
//! back from the serialised structs, so the records do not depend on this file guessing the reference's draw order;
//! only the six CRS draws are shadowed (generator.rs:86-93) because the reference does not keep them.
#![allow(non_snake_case)]

use ark_bls12_381::Bls12_381 as F;
use ark_ec::pairing::{Pairing, PairingOutput};
use ark_ec::{AffineRepr, CurveGroup};
use ark_serialize::CanonicalSerialize;
use ark_std::ops::Mul;
use ark_std::rand::{rngs::StdRng, SeedableRng};
use ark_std::UniformRand;

use groth_sahai::data_structures::*;
use groth_sahai::prover::*;
use groth_sahai::statement::*;
use groth_sahai::verifier::Verifiable;
use groth_sahai::{AbstractCrs, CRS};

type G1 = <F as Pairing>::G1;
type G2 = <F as Pairing>::G2;
type G1Affine = <F as Pairing>::G1Affine;
type G2Affine = <F as Pairing>::G2Affine;
type Fr = <F as Pairing>::ScalarField;
type GT = PairingOutput<F>;

fn ser<T: CanonicalSerialize>(t: &T) -> Vec<u8> {
    let mut v = Vec::new();
    t.serialize_compressed(&mut v).unwrap();
    v
}

fn hex(b: &[u8]) -> String {
    b.iter().map(|x| format!("{:02x}", x)).collect()
}

fn q<T: CanonicalSerialize>(t: &T) -> String {
    format!("\"{}\"", hex(&ser(t)))
}

fn list<T: CanonicalSerialize>(ts: &[T]) -> String {
    format!("[{}]", ts.iter().map(q).collect::<Vec<_>>().join(","))
}

fn u64_at(b: &[u8], off: usize) -> usize {
    let mut w = [0u8; 8];
    w.copy_from_slice(&b[off..off + 8]);
    u64::from_le_bytes(w) as usize
}

/// JSON rows of the Matrix<Fr> serialised at `off` (u64 row count, then per row u64 length + 32-byte scalars).
fn matrix_at(b: &[u8], mut off: usize) -> String {
    let rows = u64_at(b, off);
    off += 8;
    let mut out = Vec::new();
    for _ in 0..rows {
        let k = u64_at(b, off);
        off += 8;
        let mut row = Vec::new();
        for _ in 0..k {
            row.push(format!("\"{}\"", hex(&b[off..off + 32])));
            off += 32;
        }
        out.push(format!("[{}]", row.join(",")));
    }
    assert_eq!(off, b.len());
    format!("[{}]", out.join(","))
}

/// rand of a serialised Commit1 (which = 1: 96 B per commitment) or Commit2 (which = 2: 192 B)
fn commit_rand(bytes: &[u8], which: usize) -> String {
    let n = u64_at(bytes, 0);
    matrix_at(bytes, 8 + n * 96 * which)
}

/// T of a serialised EquProof: Vec<Com2> pi, Vec<Com1> theta, 1-byte EquType, Matrix<Fr>
fn proof_rand(bytes: &[u8]) -> String {
    let np = u64_at(bytes, 0);
    let off = 8 + np * 192;
    let nt = u64_at(bytes, off);
    matrix_at(bytes, off + 8 + nt * 96 + 1)
}

fn commit_record(kind: &str, crs: &CRS<F>, vars: String, commit: Vec<u8>, which: usize) {
    println!(
        "{{\"kind\":\"{}\",\"source\":\"arkworks\",\"crs\":{},\"vars\":{},\"rand\":{},\"commit\":\"{}\"}}",
        kind,
        q(crs),
        vars,
        commit_rand(&commit, which),
        hex(&commit)
    );
}

#[allow(clippy::too_many_arguments)]
fn prove_record(ty: u8, crs: &CRS<F>, equation: String, xvars: String, yvars: String, cp: &CProof<F>, ok: bool) {
    let xc = ser(&cp.xcoms);
    let yc = ser(&cp.ycoms);
    let pf = ser(&cp.equ_proofs[0]);
    println!(
        "{{\"kind\":\"prove\",\"source\":\"arkworks\",\"equ_type\":{},\"crs\":{},\"equation\":{},\"xvars\":{},\"yvars\":{},\
         \"xrand\":{},\"yrand\":{},\"T\":{},\"xcoms\":\"{}\",\"ycoms\":\"{}\",\"proof\":\"{}\",\"verify\":{}}}",
        ty,
        q(crs),
        equation,
        xvars,
        yvars,
        commit_rand(&xc, 1),
        commit_rand(&yc, 2),
        proof_rand(&pf),
        hex(&xc),
        hex(&yc),
        hex(&pf),
        ok
    );
}

fn main() {
    let mut rng = StdRng::seed_from_u64(0x6773_5f62_3230_30);

    // ---- crs: shadow the six draws p1 <- G1, p2 <- G2, a1, a2, t1, t2 <- Fr (generator.rs:86-93)
    let mut shadow = rng.clone();
    let crs = CRS::<F>::generate_crs(&mut rng);
    let p1 = G1::rand(&mut shadow).into_affine();
    let p2 = G2::rand(&mut shadow).into_affine();
    let fr: Vec<Fr> = (0..4).map(|_| Fr::rand(&mut shadow)).collect();
    println!(
        "{{\"kind\":\"crs\",\"source\":\"arkworks\",\"draws\":{{\"p1\":{},\"p2\":{},\"fr\":{}}},\"crs\":{}}}",
        q(&p1),
        q(&p2),
        list(&fr),
        q(&crs)
    );

    // ---- the four commit variants (commit.rs:78-100, 178-200, 125-156, 225-256); one identity among the group values
    let xs: Vec<G1Affine> = vec![G1::rand(&mut rng).into_affine(), G1Affine::zero(), G1::rand(&mut rng).into_affine()];
    let c = batch_commit_G1(&xs, &crs, &mut rng);
    commit_record("commit_g1", &crs, list(&xs), ser(&c), 1);
    let ys: Vec<G2Affine> = vec![G2::rand(&mut rng).into_affine(), G2Affine::zero(), G2::rand(&mut rng).into_affine()];
    let c = batch_commit_G2(&ys, &crs, &mut rng);
    commit_record("commit_g2", &crs, list(&ys), ser(&c), 2);
    let sx: Vec<Fr> = (0..3).map(|_| Fr::rand(&mut rng)).collect();
    let c = batch_commit_scalar_to_B1(&sx, &crs, &mut rng);
    commit_record("commit_b1", &crs, list(&sx), ser(&c), 1);
    let sy: Vec<Fr> = (0..3).map(|_| Fr::rand(&mut rng)).collect();
    let c = batch_commit_scalar_to_B2(&sy, &crs, &mut rng);
    commit_record("commit_b2", &crs, list(&sy), ser(&c), 2);

    // ---- one satisfied equation of every type, m = 2 x-variables, n = 1 y-variable, Gamma = [[g0], [g1]];
    //      b_consts[0] is the identity / zero so that trivial terms are covered (tests/prover.rs shapes)
    let (g0, g1) = (Fr::rand(&mut rng), Fr::from(5u64));
    let gamma: Matrix<Fr> = vec![vec![g0], vec![g1]];

    {
        // PPE: e(a0, y0) + e(x0, b0) + e(x1, b1) + g0 e(x0, y0) + g1 e(x1, y0) = t
        let x: Vec<G1Affine> = (0..2).map(|_| G1::rand(&mut rng).into_affine()).collect();
        let y: Vec<G2Affine> = vec![G2::rand(&mut rng).into_affine()];
        let a: Vec<G1Affine> = vec![G1::rand(&mut rng).into_affine()];
        let b: Vec<G2Affine> = vec![G2Affine::zero(), G2::rand(&mut rng).into_affine()];
        let t: GT = F::pairing(a[0], y[0])
            + F::pairing(x[0], b[0])
            + F::pairing(x[1], b[1])
            + F::pairing(x[0], y[0]) * g0
            + F::pairing(x[1], y[0]) * g1;
        let equ = PPE::<F> { a_consts: a, b_consts: b, gamma: gamma.clone(), target: t };
        let cp = equ.commit_and_prove(&x, &y, &crs, &mut rng);
        let ok = equ.verify(&cp, &crs);
        prove_record(0, &crs, q(&equ), list(&x), list(&y), &cp, ok);
    }
    {
        // MSMEG1: y0 a0 + b0 x0 + b1 x1 + g0 y0 x0 + g1 y0 x1 = t in G1
        let x: Vec<G1Affine> = (0..2).map(|_| G1::rand(&mut rng).into_affine()).collect();
        let y: Vec<Fr> = vec![Fr::rand(&mut rng)];
        let a: Vec<G1Affine> = vec![G1::rand(&mut rng).into_affine()];
        let b: Vec<Fr> = vec![Fr::from(0u64), Fr::rand(&mut rng)];
        let t: G1Affine = (a[0].mul(y[0]) + x[0].mul(b[0]) + x[1].mul(b[1]) + x[0].mul(g0 * y[0]) + x[1].mul(g1 * y[0]))
            .into_affine();
        let equ = MSMEG1::<F> { a_consts: a, b_consts: b, gamma: gamma.clone(), target: t };
        let cp = equ.commit_and_prove(&x, &y, &crs, &mut rng);
        let ok = equ.verify(&cp, &crs);
        prove_record(1, &crs, q(&equ), list(&x), list(&y), &cp, ok);
    }
    {
        // MSMEG2: a0 y0 + x0 b0 + x1 b1 + g0 x0 y0 + g1 x1 y0 = t in G2
        let x: Vec<Fr> = (0..2).map(|_| Fr::rand(&mut rng)).collect();
        let y: Vec<G2Affine> = vec![G2::rand(&mut rng).into_affine()];
        let a: Vec<Fr> = vec![Fr::rand(&mut rng)];
        let b: Vec<G2Affine> = vec![G2Affine::zero(), G2::rand(&mut rng).into_affine()];
        let t: G2Affine = (y[0].mul(a[0]) + b[0].mul(x[0]) + b[1].mul(x[1]) + y[0].mul(g0 * x[0]) + y[0].mul(g1 * x[1]))
            .into_affine();
        let equ = MSMEG2::<F> { a_consts: a, b_consts: b, gamma: gamma.clone(), target: t };
        let cp = equ.commit_and_prove(&x, &y, &crs, &mut rng);
        let ok = equ.verify(&cp, &crs);
        prove_record(2, &crs, q(&equ), list(&x), list(&y), &cp, ok);
    }
    {
        // Quadratic: a0 y0 + x0 b0 + x1 b1 + g0 x0 y0 + g1 x1 y0 = t in Fr
        let x: Vec<Fr> = (0..2).map(|_| Fr::rand(&mut rng)).collect();
        let y: Vec<Fr> = vec![Fr::rand(&mut rng)];
        let a: Vec<Fr> = vec![Fr::rand(&mut rng)];
        let b: Vec<Fr> = vec![Fr::from(0u64), Fr::rand(&mut rng)];
        let t: Fr = a[0] * y[0] + x[0] * b[0] + x[1] * b[1] + g0 * x[0] * y[0] + g1 * x[1] * y[0];
        let equ = QuadEqu::<F> { a_consts: a, b_consts: b, gamma: gamma.clone(), target: t };
        let cp = equ.commit_and_prove(&x, &y, &crs, &mut rng);
        let ok = equ.verify(&cp, &crs);
        prove_record(3, &crs, q(&equ), list(&x), list(&y), &cp, ok);
    }

    // ---- ComT::pairing_sum with identities on either side (data_structures.rs:494-502; entries row-major :1361-1377)
    let xs: Vec<Com1<F>> = vec![
        Com1::<F>(G1::rand(&mut rng).into_affine(), G1::rand(&mut rng).into_affine()),
        Com1::<F>(G1Affine::zero(), G1::rand(&mut rng).into_affine()),
    ];
    let ys: Vec<Com2<F>> = vec![
        Com2::<F>(G2::rand(&mut rng).into_affine(), G2::rand(&mut rng).into_affine()),
        Com2::<F>(G2::rand(&mut rng).into_affine(), G2Affine::zero()),
    ];
    let m = ComT::<F>::pairing_sum(&xs, &ys).as_matrix();
    let entries: Vec<GT> = vec![m[0][0], m[0][1], m[1][0], m[1][1]];
    println!(
        "{{\"kind\":\"pairing_sum\",\"source\":\"arkworks\",\"xs\":{},\"ys\":{},\"comt\":{}}}",
        list(&xs),
        list(&ys),
        list(&entries)
    );
}
