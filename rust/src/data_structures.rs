//! `Com1 / Com2 / ComT`, the `B / B1 / B2 / BT` traits and `Mat` (reference: `src/data_structures.rs:37-913`), with every
//! group operation, scalar multiplication, pairing and matrix product forwarded to the CUDA engine.
//! A single `a + b` is a batch of one for the GPU; code that cares about throughput should use the `Mat` / `Sum` forms
//! (one launch per matrix / slice) or the batched entry points of `prover` and `verifier`.
use crate::ffi::{self, *};
use crate::generator::CRS;
use ark_ec::{pairing::PairingOutput, AffineRepr};
use ark_ff::{Field, Zero};
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize};
use std::{fmt::Debug, iter::Sum, ops::{Add, AddAssign, Neg, Sub, SubAssign}};

pub type Matrix<T> = Vec<Vec<T>>;

/// Matrix arithmetic over field elements or commitment-group elements (`:37-46`).
pub trait Mat<Elem: Clone>: Eq + Clone + Debug {
    type Other;
    fn add(&self, other: &Self) -> Self;
    fn neg(&self) -> Self;
    fn scalar_mul(&self, other: &Self::Other) -> Self;
    fn transpose(&self) -> Self;
    fn left_mul(&self, lhs: &Matrix<Self::Other>, is_parallel: bool) -> Self;
    fn right_mul(&self, rhs: &Matrix<Self::Other>, is_parallel: bool) -> Self;
}

pub trait B<E: Gpu>: Eq + Copy + Clone + Debug + Zero + Add<Self, Output = Self> + AddAssign<Self> + Sub<Self, Output = Self>
    + SubAssign<Self> + Neg<Output = Self> + Sum {}

pub trait B1<E: Gpu>: B<E> + From<Matrix<E::G1Affine>> {
    fn as_col_vec(&self) -> Matrix<E::G1Affine>;
    fn as_vec(&self) -> Vec<E::G1Affine>;
    fn linear_map(x: &E::G1Affine) -> Self;
    fn batch_linear_map(x_vec: &[E::G1Affine]) -> Vec<Self>;
    fn scalar_linear_map(x: &E::ScalarField, key: &CRS<E>) -> Self;
    fn batch_scalar_linear_map(x_vec: &[E::ScalarField], key: &CRS<E>) -> Vec<Self>;
    fn scalar_mul(&self, other: &E::ScalarField) -> Self;
}
pub trait B2<E: Gpu>: B<E> + From<Matrix<E::G2Affine>> {
    fn as_col_vec(&self) -> Matrix<E::G2Affine>;
    fn as_vec(&self) -> Vec<E::G2Affine>;
    fn linear_map(y: &E::G2Affine) -> Self;
    fn batch_linear_map(y_vec: &[E::G2Affine]) -> Vec<Self>;
    fn scalar_linear_map(y: &E::ScalarField, key: &CRS<E>) -> Self;
    fn batch_scalar_linear_map(y_vec: &[E::ScalarField], key: &CRS<E>) -> Vec<Self>;
    fn scalar_mul(&self, other: &E::ScalarField) -> Self;
}
pub trait BT<E: Gpu, C1: B1<E>, C2: B2<E>>: B<E> + From<Matrix<PairingOutput<E>>> {
    fn as_matrix(&self) -> Matrix<PairingOutput<E>>;
    fn pairing(x: C1, y: C2) -> Self;
    fn pairing_sum(x_vec: &[C1], y_vec: &[C2]) -> Self;
    fn linear_map_PPE(z: &PairingOutput<E>) -> Self;
    fn linear_map_MSMEG1(z: &E::G1Affine, key: &CRS<E>) -> Self;
    fn linear_map_MSMEG2(z: &E::G2Affine, key: &CRS<E>) -> Self;
    fn linear_map_quad(z: &E::ScalarField, key: &CRS<E>) -> Self;
}

#[derive(Copy, Clone, Debug, CanonicalSerialize, CanonicalDeserialize)]
pub struct Com1<E: Gpu>(pub E::G1Affine, pub E::G1Affine);
#[derive(Copy, Clone, Debug, CanonicalSerialize, CanonicalDeserialize)]
pub struct Com2<E: Gpu>(pub E::G2Affine, pub E::G2Affine);
#[derive(Copy, Clone, Debug)]
pub struct ComT<E: Gpu>(pub PairingOutput<E>, pub PairingOutput<E>, pub PairingOutput<E>, pub PairingOutput<E>);

pub fn col_vec_to_vec<F: Clone>(mat: &Matrix<F>) -> Vec<F> {
    if mat.len() == 1 { mat[0].clone() } else { mat.iter().map(|row| row[0].clone()).collect() }     // `:145-151`
}
pub fn vec_to_col_vec<F: Clone>(vec: &[F]) -> Matrix<F> { vec.iter().map(|e| vec![e.clone()]).collect() }

// ---------------------------------------------------------------- ABI views
impl<E: Gpu> Com1<E> {
    pub(crate) fn abi(&self) -> GsCom1 { GsCom1([E::g1(&self.0), E::g1(&self.1)]) }
    pub(crate) fn from_abi(c: &GsCom1) -> Self { Com1(E::g1_back(&c.0[0]), E::g1_back(&c.0[1])) }
}
impl<E: Gpu> Com2<E> {
    pub(crate) fn abi(&self) -> GsCom2 { GsCom2([E::g2(&self.0), E::g2(&self.1)]) }
    pub(crate) fn from_abi(c: &GsCom2) -> Self { Com2(E::g2_back(&c.0[0]), E::g2_back(&c.0[1])) }
}
impl<E: Gpu> ComT<E> {
    pub(crate) fn abi(&self) -> GsComT { GsComT([E::gt(&self.0), E::gt(&self.1), E::gt(&self.2), E::gt(&self.3)]) }
    pub(crate) fn from_abi(c: &GsComT) -> Self { ComT(E::gt_back(&c.0[0]), E::gt_back(&c.0[1]), E::gt_back(&c.0[2]), E::gt_back(&c.0[3])) }
}
pub(crate) fn com1s<E: Gpu>(v: &[Com1<E>]) -> Vec<GsCom1> { v.iter().map(|c| c.abi()).collect() }
pub(crate) fn com2s<E: Gpu>(v: &[Com2<E>]) -> Vec<GsCom2> { v.iter().map(|c| c.abi()).collect() }

// ---------------------------------------------------------------- Add / Sub / Neg / Sum / Zero / Eq  (`:162-278`, `:391-479`)
macro_rules! group_ops {
    ($com:ident, $abi:ident, $add:ident, $sub:ident, $neg:ident, $sum:ident) => {
        impl<E: Gpu> PartialEq for $com<E> { fn eq(&self, o: &Self) -> bool { self.parts() == o.parts() } }
        impl<E: Gpu> Eq for $com<E> {}
        impl<E: Gpu> Add for $com<E> {
            type Output = Self;
            fn add(self, o: Self) -> Self {
                let (a, b, mut r) = (self.abi(), o.abi(), self.abi());
                with_ctx(|c| check(c, unsafe { ffi::$add(c.raw(), 1, &a, &b, &mut r) }));
                Self::from_abi(&r)
            }
        }
        impl<E: Gpu> Sub for $com<E> {
            type Output = Self;
            fn sub(self, o: Self) -> Self {
                let (a, b, mut r) = (self.abi(), o.abi(), self.abi());
                with_ctx(|c| check(c, unsafe { ffi::$sub(c.raw(), 1, &a, &b, &mut r) }));
                Self::from_abi(&r)
            }
        }
        impl<E: Gpu> Neg for $com<E> {
            type Output = Self;
            fn neg(self) -> Self {
                let (a, mut r) = (self.abi(), self.abi());
                with_ctx(|c| check(c, unsafe { ffi::$neg(c.raw(), 1, &a, &mut r) }));
                Self::from_abi(&r)
            }
        }
        impl<E: Gpu> AddAssign for $com<E> { fn add_assign(&mut self, o: Self) { *self = *self + o; } }
        impl<E: Gpu> SubAssign for $com<E> { fn sub_assign(&mut self, o: Self) { *self = *self - o; } }
        impl<E: Gpu> Sum for $com<E> {
            /// One launch for the whole iterator (the reference folds `a + b`, normalising after every term).
            fn sum<I: Iterator<Item = Self>>(iter: I) -> Self {
                let v: Vec<$abi> = iter.map(|x| x.abi()).collect();
                let mut r = Self::zero().abi();
                with_ctx(|c| check(c, unsafe { ffi::$sum(c.raw(), v.len(), v.as_ptr(), &mut r) }));
                Self::from_abi(&r)
            }
        }
        impl<E: Gpu> B<E> for $com<E> {}
    };
}
impl<E: Gpu> Com1<E> { fn parts(&self) -> (E::G1Affine, E::G1Affine) { (self.0, self.1) } }
impl<E: Gpu> Com2<E> { fn parts(&self) -> (E::G2Affine, E::G2Affine) { (self.0, self.1) } }
impl<E: Gpu> ComT<E> { fn parts(&self) -> [PairingOutput<E>; 4] { [self.0, self.1, self.2, self.3] } }
group_ops!(Com1, GsCom1, gs_com1_add, gs_com1_sub, gs_com1_neg, gs_com1_sum);
group_ops!(Com2, GsCom2, gs_com2_add, gs_com2_sub, gs_com2_neg, gs_com2_sum);
group_ops!(ComT, GsComT, gs_comt_add, gs_comt_sub, gs_comt_neg, gs_comt_sum);

impl<E: Gpu> Zero for Com1<E> {
    fn zero() -> Self { Com1(E::G1Affine::zero(), E::G1Affine::zero()) }
    fn is_zero(&self) -> bool { *self == Self::zero() }
}
impl<E: Gpu> Zero for Com2<E> {
    fn zero() -> Self { Com2(E::G2Affine::zero(), E::G2Affine::zero()) }
    fn is_zero(&self) -> bool { *self == Self::zero() }
}
impl<E: Gpu> Zero for ComT<E> {       // GT is written additively by arkworks: zero = the Fp12 one
    fn zero() -> Self { let z = PairingOutput::<E>::zero(); ComT(z, z, z, z) }
    fn is_zero(&self) -> bool { *self == Self::zero() }
}
impl<E: Gpu> From<Matrix<E::G1Affine>> for Com1<E> {
    fn from(mat: Matrix<E::G1Affine>) -> Self { assert_eq!((mat.len(), mat[0].len()), (2, 1)); Com1(mat[0][0], mat[1][0]) }
}
impl<E: Gpu> From<Matrix<E::G2Affine>> for Com2<E> {
    fn from(mat: Matrix<E::G2Affine>) -> Self { assert_eq!((mat.len(), mat[0].len()), (2, 1)); Com2(mat[0][0], mat[1][0]) }
}
impl<E: Gpu> From<Matrix<PairingOutput<E>>> for ComT<E> {
    fn from(mat: Matrix<PairingOutput<E>>) -> Self {
        assert_eq!((mat.len(), mat[0].len(), mat[1].len()), (2, 2, 2));
        ComT(mat[0][0], mat[0][1], mat[1][0], mat[1][1])
    }
}

// ---------------------------------------------------------------- B1 / B2: linear maps and scalar multiplication (`:300-387`)
impl<E: Gpu> B1<E> for Com1<E> {
    fn as_col_vec(&self) -> Matrix<E::G1Affine> { vec![vec![self.0], vec![self.1]] }
    fn as_vec(&self) -> Vec<E::G1Affine> { vec![self.0, self.1] }
    fn linear_map(x: &E::G1Affine) -> Self { Com1(E::G1Affine::zero(), *x) }                          // iota_1
    fn batch_linear_map(x_vec: &[E::G1Affine]) -> Vec<Self> { x_vec.iter().map(Self::linear_map).collect() }
    fn scalar_linear_map(x: &E::ScalarField, key: &CRS<E>) -> Self { Self::batch_scalar_linear_map(&[*x], key)[0] }
    /// iota_1'(x) = x * W1 with W1 = u2 + (O, g1): one fixed-base call for the whole slice (commitment with zero
    /// randomness: `gs_batch_commit_scalar_b1(xs, 0)` = x W1 + 0 u1).
    fn batch_scalar_linear_map(x_vec: &[E::ScalarField], key: &CRS<E>) -> Vec<Self> {
        let (xs, zeros) = (frs::<E>(x_vec), vec![GsFr::default(); x_vec.len()]);
        let mut out = vec![Self::zero().abi(); x_vec.len()];
        with_crs(&key.abi(), |c| check(c, unsafe { gs_batch_commit_scalar_b1(c.raw(), xs.len(), xs.as_ptr(), zeros.as_ptr(), out.as_mut_ptr()) }));
        out.iter().map(Self::from_abi).collect()
    }
    fn scalar_mul(&self, other: &E::ScalarField) -> Self {                                            // `:336-342`
        let (s, a, mut r) = (E::fr(other), self.abi(), self.abi());
        with_ctx(|c| check(c, unsafe { gs_com1_matmul(c.raw(), 1, 1, 1, &s, &a, &mut r) }));
        Self::from_abi(&r)
    }
}
impl<E: Gpu> B2<E> for Com2<E> {
    fn as_col_vec(&self) -> Matrix<E::G2Affine> { vec![vec![self.0], vec![self.1]] }
    fn as_vec(&self) -> Vec<E::G2Affine> { vec![self.0, self.1] }
    fn linear_map(y: &E::G2Affine) -> Self { Com2(E::G2Affine::zero(), *y) }                          // iota_2
    fn batch_linear_map(y_vec: &[E::G2Affine]) -> Vec<Self> { y_vec.iter().map(Self::linear_map).collect() }
    fn scalar_linear_map(y: &E::ScalarField, key: &CRS<E>) -> Self { Self::batch_scalar_linear_map(&[*y], key)[0] }
    fn batch_scalar_linear_map(y_vec: &[E::ScalarField], key: &CRS<E>) -> Vec<Self> {
        let (ys, zeros) = (frs::<E>(y_vec), vec![GsFr::default(); y_vec.len()]);
        let mut out = vec![Self::zero().abi(); y_vec.len()];
        with_crs(&key.abi(), |c| check(c, unsafe { gs_batch_commit_scalar_b2(c.raw(), ys.len(), ys.as_ptr(), zeros.as_ptr(), out.as_mut_ptr()) }));
        out.iter().map(Self::from_abi).collect()
    }
    fn scalar_mul(&self, other: &E::ScalarField) -> Self {                                            // `:381-387`
        let (s, a, mut r) = (E::fr(other), self.abi(), self.abi());
        with_ctx(|c| check(c, unsafe { gs_com2_matmul(c.raw(), 1, 1, 1, &s, &a, &mut r) }));
        Self::from_abi(&r)
    }
}

// ---------------------------------------------------------------- BT: pairing, pairing_sum, iota_T (`:484-540`)
impl<E: Gpu> ComT<E> {
    fn linear_map_typed(ty: i32, target: *const u8, key: Option<&CRS<E>>) -> Self {
        let mut out = Self::zero().abi();
        let run = |c: &Ctx| check(c, unsafe { gs_comt_linear_map(c.raw(), ty, target, &mut out) });
        match key { Some(k) => with_crs(&k.abi(), run), None => with_ctx(run) }
        Self::from_abi(&out)
    }
}
impl<E: Gpu> BT<E, Com1<E>, Com2<E>> for ComT<E> {
    fn as_matrix(&self) -> Matrix<PairingOutput<E>> { vec![vec![self.0, self.1], vec![self.2, self.3]] }
    /// Row-major `[e(x0,y0), e(x0,y1), e(x1,y0), e(x1,y1)]`.
    fn pairing(x: Com1<E>, y: Com2<E>) -> Self {
        let (a, b, mut out) = (x.abi(), y.abi(), Self::zero().abi());
        with_ctx(|c| check(c, unsafe { gs_comt_pairing(c.raw(), 1, &a, &b, &mut out) }));
        Self::from_abi(&out)
    }
    /// One Miller accumulator and ONE final exponentiation per entry for the whole slice.
    fn pairing_sum(x_vec: &[Com1<E>], y_vec: &[Com2<E>]) -> Self {
        assert_eq!(x_vec.len(), y_vec.len());                                                         // `:495`
        let (xs, ys, mut out) = (com1s(x_vec), com2s(y_vec), Self::zero().abi());
        with_ctx(|c| check(c, unsafe { gs_comt_pairing_sum(c.raw(), xs.len(), xs.as_ptr(), ys.as_ptr(), &mut out) }));
        Self::from_abi(&out)
    }
    fn linear_map_PPE(z: &PairingOutput<E>) -> Self { let o = PairingOutput::<E>::zero(); ComT(o, o, o, *z) }
    fn linear_map_MSMEG1(z: &E::G1Affine, key: &CRS<E>) -> Self { let t = E::g1(z); Self::linear_map_typed(1, &t as *const _ as *const u8, Some(key)) }
    fn linear_map_MSMEG2(z: &E::G2Affine, key: &CRS<E>) -> Self { let t = E::g2(z); Self::linear_map_typed(2, &t as *const _ as *const u8, Some(key)) }
    fn linear_map_quad(z: &E::ScalarField, key: &CRS<E>) -> Self { let t = E::fr(z); Self::linear_map_typed(3, &t as *const _ as *const u8, Some(key)) }
}

// ---------------------------------------------------------------- Mat on Matrix<Com1 / Com2> (`:545-742`)
fn dims<T>(m: &Matrix<T>) -> (usize, usize) { (m.len(), m.first().map_or(0, |r| r.len())) }
fn transpose_of<T: Clone>(m: &Matrix<T>) -> Matrix<T> {
    let (r, c) = dims(m);
    (0..c).map(|j| (0..r).map(|i| m[i][j].clone()).collect()).collect()
}
macro_rules! com_mat {
    ($com:ident, $abi:ident, $add:ident, $neg:ident, $matmul:ident) => {
        impl<E: Gpu> Mat<$com<E>> for Matrix<$com<E>> {
            type Other = E::ScalarField;
            fn add(&self, other: &Self) -> Self {
                assert_eq!(dims(self), dims(other));                                                 // `:591-592`
                let (r, c) = dims(self);
                let a: Vec<$abi> = self.iter().flatten().map(|x| x.abi()).collect();
                let b: Vec<$abi> = other.iter().flatten().map(|x| x.abi()).collect();
                let mut o = a.clone();
                with_ctx(|cx| check(cx, unsafe { ffi::$add(cx.raw(), a.len(), a.as_ptr(), b.as_ptr(), o.as_mut_ptr()) }));
                (0..r).map(|i| (0..c).map(|j| $com::from_abi(&o[i * c + j])).collect()).collect()
            }
            fn neg(&self) -> Self {
                let (r, c) = dims(self);
                let a: Vec<$abi> = self.iter().flatten().map(|x| x.abi()).collect();
                let mut o = a.clone();
                with_ctx(|cx| check(cx, unsafe { ffi::$neg(cx.raw(), a.len(), a.as_ptr(), o.as_mut_ptr()) }));
                (0..r).map(|i| (0..c).map(|j| $com::from_abi(&o[i * c + j])).collect()).collect()
            }
            /// Every entry times the same scalar: a (r*c x 1) * (1 x 1) product per entry would be r*c launches, so the
            /// entries are laid out as a k = 1 product with lhs = [s; r*c] rows against each entry (one call per entry kept
            /// simple here; proofs never call this on anything bigger than 2 x 1).
            fn scalar_mul(&self, other: &Self::Other) -> Self {
                let s = E::fr(other);
                self.iter().map(|row| row.iter().map(|e| {
                    let (a, mut o) = (e.abi(), e.abi());
                    with_ctx(|cx| check(cx, unsafe { ffi::$matmul(cx.raw(), 1, 1, 1, &s, &a, &mut o) }));
                    $com::from_abi(&o)
                }).collect()).collect()
            }
            fn transpose(&self) -> Self { transpose_of(self) }
            /// out (rows x c) = lhs (rows x k, Fr) * self (k x c): `is_parallel` is meaningless on the GPU (`:696-742`).
            fn left_mul(&self, lhs: &Matrix<Self::Other>, _is_parallel: bool) -> Self {
                if self.is_empty() || self[0].is_empty() || lhs.is_empty() || lhs[0].is_empty() { return vec![]; }   // `:697-702`
                let ((k, c), (rows, k2)) = (dims(self), dims(lhs));
                assert_eq!(k, k2);                                                                   // `:705`
                let (l, m) = (fr_matrix::<E>(lhs), self.iter().flatten().map(|x| x.abi()).collect::<Vec<$abi>>());
                let mut o = vec![$com::<E>::zero().abi(); rows * c];
                with_ctx(|cx| check(cx, unsafe { ffi::$matmul(cx.raw(), rows, k, c, l.as_ptr(), m.as_ptr(), o.as_mut_ptr()) }));
                (0..rows).map(|i| (0..c).map(|j| $com::from_abi(&o[i * c + j])).collect()).collect()
            }
            /// out (r x c) = self (r x k) * rhs (k x c, Fr) = (rhs^T * self^T)^T through the same kernel (`:645-694`).
            fn right_mul(&self, rhs: &Matrix<Self::Other>, is_parallel: bool) -> Self {
                if self.is_empty() || self[0].is_empty() || rhs.is_empty() || rhs[0].is_empty() { return vec![]; }
                assert_eq!(dims(self).1, dims(rhs).0);                                               // `:654`
                transpose_of(&transpose_of(self).left_mul(&transpose_of(rhs), is_parallel))
            }
        }
    };
}
com_mat!(Com1, GsCom1, gs_com1_add, gs_com1_neg, gs_com1_matmul);
com_mat!(Com2, GsCom2, gs_com2_add, gs_com2_neg, gs_com2_matmul);

// ---------------------------------------------------------------- Mat on Matrix<Fr> (`:768-913`)
/// Implemented for the scalar field of the supported curve (the reference is generic over `F: Field`; the field
/// matrices of a proof are tiny next to the group work, the big one -- R^T Gamma -- is fused into `gs_prove`).
impl Mat<ark_bls12_381::Fr> for Matrix<ark_bls12_381::Fr> {
    type Other = ark_bls12_381::Fr;
    fn add(&self, other: &Self) -> Self {
        assert_eq!(dims(self), dims(other));
        self.iter().zip(other).map(|(a, b)| a.iter().zip(b).map(|(x, y)| *x + *y).collect()).collect()
    }
    fn neg(&self) -> Self { self.iter().map(|r| r.iter().map(|x| -*x).collect()).collect() }
    fn scalar_mul(&self, other: &Self::Other) -> Self { self.iter().map(|r| r.iter().map(|x| *x * *other).collect()).collect() }
    fn transpose(&self) -> Self { transpose_of(self) }
    fn right_mul(&self, rhs: &Matrix<Self::Other>, _is_parallel: bool) -> Self {
        if self.is_empty() || self[0].is_empty() || rhs.is_empty() || rhs[0].is_empty() { return vec![]; }
        let ((r, k), (k2, c)) = (dims(self), dims(rhs));
        assert_eq!(k, k2);
        type E = ark_bls12_381::Bls12_381;
        let (a, b) = (fr_matrix::<E>(self), fr_matrix::<E>(rhs));
        let mut o = vec![GsFr::default(); r * c];
        with_ctx(|cx| check(cx, unsafe { gs_fr_matmul(cx.raw(), r, k, c, a.as_ptr(), b.as_ptr(), o.as_mut_ptr()) }));
        (0..r).map(|i| (0..c).map(|j| <E as Gpu>::fr_back(&o[i * c + j])).collect()).collect()
    }
    fn left_mul(&self, lhs: &Matrix<Self::Other>, is_parallel: bool) -> Self { lhs.right_mul(self, is_parallel) }
}

#[allow(dead_code)]
fn _assert_field<F: Field>() {}
