//! `GpuSteepestEdge`: relp's `PivotRule` (strategy/pivot_rule.rs:23-54) with pricing, the steepest-edge
//! weights and their Goldfarb-Reid update on the device (`rg_rule_new`, `rg_select_primal_pivot_column`, the
//! `update_rule` flag of `rg_bring_into_basis`).
//!
//! `PivotRule` methods are generic over `IM`, and `Tableau.inverse_maintainer` is a private field
//! (tableau/mod.rs:29).  relp already enables `#![feature(specialization)]` (src/lib.rs:8); this rule needs the
//! one-line additive accessor `Tableau::inverse_maintainer(&self) -> &IM` and is only meaningful for
//! `IM = GpuCarry` (checked at run time through `Any`).
use std::any::Any;

use relp::algorithm::two_phase::matrix_provider::column::Column;
use relp::algorithm::two_phase::strategy::pivot_rule::PivotRule;
use relp::algorithm::two_phase::tableau::inverse_maintenance::{ops as im_ops, InverseMaintainer};
use relp::algorithm::two_phase::tableau::kind::Kind;
use relp::algorithm::two_phase::tableau::{BasisChangeComputationInfo, Tableau};
use relp::data::linear_algebra::SparseTuple;
use relp_num::RationalBig;

use crate::carry::GpuCarry;
use crate::ffi;

pub struct GpuSteepestEdge;

fn gpu<IM: InverseMaintainer + 'static, K: Kind>(tableau: &Tableau<IM, K>) -> &GpuCarry {
    (tableau.inverse_maintainer() as &dyn Any).downcast_ref::<GpuCarry>()
        .expect("GpuSteepestEdge drives a GpuCarry")
}

impl PivotRule<RationalBig> for GpuSteepestEdge {
    // pivot_rule.rs:202-219 + initial_gamma :299-305, on the device
    fn new<IM, K>(tableau: &Tableau<IM, K>) -> Self
    where IM: InverseMaintainer<F = RationalBig>, K: Kind,
          RationalBig: im_ops::Column<<K::Column as Column>::F> + im_ops::Cost<K::Cost> {
        let carry = gpu(tableau);
        // first rule of a solve: materialise the provider's columns on the device (see carry.rs)
        let na = tableau.nr_artificial_variables();
        let columns: Vec<_> = (na..tableau.nr_columns()).map(|j| tableau.original_column(j)).collect();
        carry.attach_provider(&columns, None);
        let rc = unsafe { ffi::rg_rule_new(carry.ctx(), ffi::RG_RULE_STEEPEST_EDGE) };
        assert_eq!(rc, ffi::RG_OK, "rg_rule_new");
        Self
    }

    // pivot_rule.rs:221-241: max cost^2 / gamma among negative relative costs, highest index on ties
    fn select_primal_pivot_column<IM, K>(&mut self, tableau: &Tableau<IM, K>) -> Option<SparseTuple<IM::F>>
    where IM: InverseMaintainer<F = RationalBig>, K: Kind,
          RationalBig: im_ops::Column<<K::Column as Column>::F> + im_ops::Cost<K::Cost> {
        let carry = gpu(tableau);
        let (mut status, mut q) = (0i32, 0i32);
        let rc = unsafe { ffi::rg_select_primal_pivot_column(carry.ctx(), &mut status, &mut q) };
        assert_eq!(rc, ffi::RG_OK, "rg_select_primal_pivot_column");
        if status != ffi::RG_STEP_PIVOTED { return None; }
        let j = q as usize + tableau.nr_artificial_variables();
        Some((j, tableau.relative_cost(j)))
    }

    // pivot_rule.rs:243-296: done inside rg_bring_into_basis(update_rule = 1); the weights never leave HBM
    fn after_basis_update<IM, K>(&mut self, _info: BasisChangeComputationInfo<IM::F>, _tableau: &Tableau<IM, K>)
    where IM: InverseMaintainer<F = RationalBig>, K: Kind,
          RationalBig: im_ops::Column<<K::Column as Column>::F> + im_ops::Cost<K::Cost> {}
}
