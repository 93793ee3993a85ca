//! # relp-gpu
//!
//! Drop-in `InverseMaintainer` + `PivotRule` for `relp` backed by the B200 engine `librelp_gpu.so`
//! (`include/relp_gpu.h`).  Usage, next to the reference's own call shape
//! (`data.solve_relaxation::<Carry<RationalBig, LUDecomposition<RationalBig>>>()`, tests/netlib/mod.rs:62):
//!
//! ```ignore
//! use relp_gpu::{GpuCarry, solve_relaxation_gpu};
//! let result = solve_relaxation_gpu(&matrix_data);          // OptimizationResult<RationalBig>
//! ```
//!
//! `solve_relaxation_gpu` is `two_phase::solve_relaxation` (two_phase/mod.rs:25-109) with
//! `IM = GpuCarry` and `PR = GpuSteepestEdge`; relp hard-codes `SteepestDescentAlongObjective` there
//! (two_phase/mod.rs:57,68,107) and keeps `phase_one::primal` crate-private (phase_one.rs:123), so the twin
//! function below restates those 60 lines of control flow against relp's PUBLIC tableau API.  Nothing in
//! relp's signatures changes.
//!
//! This crate is NOT compiled in the engine's repository (no Rust toolchain in its build image); see
//! Cargo.toml.  The C++ host driver `relp_b200/csrc/host/relp_host.cpp` runs the same call sequence under
//! test, and `tests/test_gpu_boundary.py` checks every getter used here against the oracle after every pivot.
#![feature(trait_alias)]

pub mod ffi;
pub mod number;
mod carry;
mod rule;

pub use carry::{GpuCarry, GpuColumn};
pub use rule::GpuSteepestEdge;
