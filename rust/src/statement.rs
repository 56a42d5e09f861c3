//! Equation types (reference: `src/statement.rs:42-192`) -- plain data kept on the host; flattened at the FFI.
//!   sum_j A_j * Y_j  +  sum_i X_i * B_i  +  sum_ij Gamma_ij (X_i * Y_j)  =  target
//! with `*` = pairing (PPE), scalar multiplication (MSMEG1 / MSMEG2) or field product (QuadEqu).
use crate::data_structures::Matrix;
use crate::ffi::Gpu;
use crate::prover::Provable;
use crate::verifier::Verifiable;
use ark_ec::pairing::PairingOutput;
use ark_serialize::{CanonicalDeserialize, CanonicalSerialize, Compress, Read, SerializationError, Valid, Validate, Write};

/// Wire tag of an equation type: ONE byte, 0..3 in this order (`statement.rs:68-73`); also the `type` argument of the C ABI.
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(u8)]
pub enum EquType { PairingProduct = 0, MultiScalarG1 = 1, MultiScalarG2 = 2, Quadratic = 3 }

impl EquType {
    pub fn abi(self) -> i32 { self as u8 as i32 }
}
impl Valid for EquType { fn check(&self) -> Result<(), SerializationError> { Ok(()) } }
impl CanonicalSerialize for EquType {
    fn serialize_with_mode<W: Write>(&self, mut w: W, _c: Compress) -> Result<(), SerializationError> {
        w.write_all(&[*self as u8]).map_err(SerializationError::IoError)
    }
    fn serialized_size(&self, _c: Compress) -> usize { 1 }
}
impl CanonicalDeserialize for EquType {
    fn deserialize_with_mode<R: Read>(mut r: R, _c: Compress, _v: Validate) -> Result<Self, SerializationError> {
        let mut b = [0u8; 1];
        r.read_exact(&mut b).map_err(SerializationError::IoError)?;
        match b[0] {
            0 => Ok(EquType::PairingProduct), 1 => Ok(EquType::MultiScalarG1), 2 => Ok(EquType::MultiScalarG2),
            3 => Ok(EquType::Quadratic), _ => Err(SerializationError::InvalidData),
        }
    }
}

pub trait Equ {}
pub trait Equation<E: Gpu, A1, A2, AT>: Equ + Provable<E, A1, A2, AT> + Verifiable<E> { fn get_type(&self) -> EquType; }

macro_rules! equation {
    ($name:ident, $a:ty, $b:ty, $t:ty, $tag:expr) => {
        #[derive(Clone, Debug, PartialEq, Eq, CanonicalSerialize, CanonicalDeserialize)]
        pub struct $name<E: Gpu> {
            /// pairs with the y-variables: length n
            pub a_consts: Vec<$a>,
            /// pairs with the x-variables: length m
            pub b_consts: Vec<$b>,
            /// m x n
            pub gamma: Matrix<E::ScalarField>,
            pub target: $t,
        }
        impl<E: Gpu> Equ for $name<E> {}
        impl<E: Gpu> Equation<E, $a, $b, $t> for $name<E> { fn get_type(&self) -> EquType { $tag } }
    };
}
equation!(PPE, E::G1Affine, E::G2Affine, PairingOutput<E>, EquType::PairingProduct);
equation!(MSMEG1, E::G1Affine, E::ScalarField, E::G1Affine, EquType::MultiScalarG1);
equation!(MSMEG2, E::ScalarField, E::G2Affine, E::G2Affine, EquType::MultiScalarG2);
equation!(QuadEqu, E::ScalarField, E::ScalarField, E::ScalarField, EquType::Quadratic);
