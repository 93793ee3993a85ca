//! Raw bindings of `include/relp_gpu.h` (what `bindgen --allowlist-function 'rg_.*'` emits; kept in the
//! tree so that the crate builds without libclang).  One declaration per entry point of the header.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct rg_context { _private: [u8; 0] }

pub const RG_OK: c_int = 0;
pub const RG_ERR_CUDA: c_int = -1;
pub const RG_ERR_ARG: c_int = -2;
pub const RG_ERR_OVERFLOW: c_int = -3;
pub const RG_ERR_STATE: c_int = -4;
pub const RG_ERR_NCCL: c_int = -5;

pub const RG_STEP_PIVOTED: i32 = 0;
pub const RG_STEP_OPTIMAL: i32 = 1;
pub const RG_STEP_UNBOUNDED: i32 = 2;

pub const RG_RULE_FIRST_PROFITABLE: i32 = 0;
pub const RG_RULE_FIRST_PROFITABLE_WITH_MEMORY: i32 = 1;
pub const RG_RULE_DANTZIG: i32 = 2;
pub const RG_RULE_STEEPEST_EDGE: i32 = 3;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct rg_options {
    pub device: i32,
    pub initial_limbs: i32,
    pub rank: i32,
    pub world: i32,
    pub dense_carry: i32,
    pub reserved: i32,
    pub nccl_unique_id: *const c_void,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct rg_pivot_info {
    pub status: i32,
    pub entering: i32,
    pub row: i32,
    pub leaving: i32,
}

unsafe extern "C" {
    pub fn rg_create(opts: *const rg_options, out: *mut *mut rg_context) -> c_int;
    pub fn rg_destroy(ctx: *mut rg_context) -> c_int;
    pub fn rg_release_cached_memory(device: i32) -> i64;
    pub fn rg_last_error(ctx: *const rg_context) -> *const c_char;
    pub fn rg_nccl_unique_id(out: *mut c_void, bytes: i32) -> c_int;
    pub fn rg_load_csc(ctx: *mut rg_context, m: i32, n: i32, colptr: *const i64, rowidx: *const i32,
                       vals: *const i64) -> c_int;
    pub fn rg_load_dense_i8(ctx: *mut rg_context, nd: i32, colmajor: *const i8) -> c_int;
    pub fn rg_set_rhs(ctx: *mut rg_context, b: *const i64) -> c_int;
    pub fn rg_set_weights(ctx: *mut rg_context, colfac: *const i64, artfac: *const i64, colw: *const i64,
                          artcost: *const i64) -> c_int;
    pub fn rg_init_identity_basis(ctx: *mut rg_context, basis: *const i32, cost: *const i64) -> c_int;
    pub fn rg_init_basis(ctx: *mut rg_context, basis: *const i32, cost: *const i64) -> c_int;
    pub fn rg_phase_switch(ctx: *mut rg_context, cost: *const i64) -> c_int;
    pub fn rg_rule_new(ctx: *mut rg_context, rule: i32) -> c_int;
    pub fn rg_select_primal_pivot_column(ctx: *mut rg_context, status: *mut i32, q: *mut i32) -> c_int;
    pub fn rg_generate_column(ctx: *mut rg_context, q: i32) -> c_int;
    pub fn rg_select_primal_pivot_row(ctx: *mut rg_context, status: *mut i32, row: *mut i32) -> c_int;
    pub fn rg_bring_into_basis(ctx: *mut rg_context, q: i32, row: i32, update_rule: i32,
                               info: *mut rg_pivot_info) -> c_int;
    pub fn rg_iterate(ctx: *mut rg_context, max_pivots: i64, trace: *mut rg_pivot_info, n_done: *mut i64,
                      status: *mut i32) -> c_int;
    pub fn rg_remove_artificial_row(ctx: *mut rg_context, row: i32, info: *mut rg_pivot_info) -> c_int;
    pub fn rg_get_limbs(ctx: *mut rg_context, limbs: *mut i32) -> c_int;
    pub fn rg_get_denominator(ctx: *mut rg_context, out: *mut u64) -> c_int;
    pub fn rg_get_basis(ctx: *mut rg_context, basis: *mut i32) -> c_int;
    pub fn rg_get_b(ctx: *mut rg_context, out: *mut u64) -> c_int;
    pub fn rg_get_minus_objective(ctx: *mut rg_context, out: *mut u64) -> c_int;
    pub fn rg_get_minus_pi(ctx: *mut rg_context, out: *mut u64) -> c_int;
    pub fn rg_get_basis_inverse_row(ctx: *mut rg_context, row: i32, out: *mut u64) -> c_int;
    pub fn rg_get_pivot_column(ctx: *mut rg_context, out: *mut u64) -> c_int;
    pub fn rg_get_relative_costs(ctx: *mut rg_context, out: *mut u64) -> c_int;
    pub fn rg_get_gamma(ctx: *mut rg_context, out: *mut u64) -> c_int;
    pub fn rg_get_element(ctx: *mut rg_context, row: i32, j: i32, out: *mut u64) -> c_int;
    pub fn rg_get_basis_change_info(ctx: *mut rg_context, column: *mut u64, work: *mut u64, row: *mut u64,
                                    denominator_before: *mut u64) -> c_int;
}
