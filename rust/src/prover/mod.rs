//! Commit and prove (reference: `src/prover/mod.rs:1-5`).
pub mod commit;
pub mod prove;
pub use commit::*;
pub use prove::*;
