//! CRS and its generation (reference: `src/generator.rs:25-119`).
use crate::data_structures::{Com1, Com2};
use crate::ffi::*;
use ark_ec::pairing::PairingOutput;
use ark_ff::UniformRand;
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize};
use ark_std::rand::Rng;

pub trait AbstractCrs<E: Gpu> { fn generate_crs<R: Rng>(rng: &mut R) -> Self; }

#[derive(Clone, Debug, CanonicalSerialize, CanonicalDeserialize)]
pub struct CRS<E: Gpu> {
    pub u: Vec<Com1<E>>,
    pub v: Vec<Com2<E>>,
    pub g1_gen: E::G1Affine,
    pub g2_gen: E::G2Affine,
    pub gt_gen: PairingOutput<E>,
}

impl<E: Gpu> CRS<E> {
    pub(crate) fn abi(&self) -> GsCrs {
        assert_eq!((self.u.len(), self.v.len()), (2, 2));
        GsCrs { u: [self.u[0].abi(), self.u[1].abi()], v: [self.v[0].abi(), self.v[1].abi()],
                g1_gen: E::g1(&self.g1_gen), g2_gen: E::g2(&self.g2_gen), gt_gen: E::gt(&self.gt_gen) }
    }
    pub(crate) fn from_abi(c: &GsCrs) -> Self {
        CRS { u: c.u.iter().map(Com1::from_abi).collect(), v: c.v.iter().map(Com2::from_abi).collect(),
              g1_gen: E::g1_back(&c.g1_gen), g2_gen: E::g2_back(&c.g2_gen), gt_gen: E::gt_back(&c.gt_gen) }
    }
}

impl<E: Gpu> AbstractCrs<E> for CRS<E> {
    /// Binding key.  The six values are drawn on the host in the reference's order -- p1 <- G1, p2 <- G2, a1, a2, t1, t2
    /// <- Fr (`generator.rs:86-93`) -- and handed to `gs_crs_generate`: u = [(p1, a1 p1), (t1 p1, t1 a1 p1)], v likewise,
    /// gt_gen = e(p1, p2); the result also becomes the thread context's current key.
    fn generate_crs<R: Rng>(rng: &mut R) -> Self {
        let p1 = E::G1::rand(rng).into();
        let p2 = E::G2::rand(rng).into();
        let (a1, a2) = (E::ScalarField::rand(rng), E::ScalarField::rand(rng));
        let (t1, t2) = (E::ScalarField::rand(rng), E::ScalarField::rand(rng));
        let (gp1, gp2) = (E::g1(&p1), E::g2(&p2));
        let s = [E::fr(&a1), E::fr(&a2), E::fr(&t1), E::fr(&t2)];
        let mut out = std::mem::MaybeUninit::<GsCrs>::zeroed();
        with_ctx(|c| check(c, unsafe { gs_crs_generate(c.raw(), &gp1, &gp2, &s[0], &s[1], &s[2], &s[3], out.as_mut_ptr()) }));
        Self::from_abi(&unsafe { out.assume_init() })
    }
}
