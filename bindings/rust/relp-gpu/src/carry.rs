//! `GpuCarry`: relp's `InverseMaintainer` (tableau/inverse_maintenance/mod.rs:30-264) on the device.
//!
//! The whole carry `[-obj | -pi ; b | B^-1]` lives in HBM as integer numerators over one denominator; this
//! type owns the `rg_context`, the host integer prescale of the problem and a host mirror of `b`
//! (`get_constraint_value` returns a reference, mod.rs:255).
//!
//! relp hands the constraint matrix to an inverse maintainer one column at a time (`generate_column(column)`),
//! never as a whole, and the artificial constructors only receive `b`.  The device wants the matrix once
//! (legal: a `MatrixProvider` is immutable during a solve, matrix_provider/mod.rs:19-22), so the device
//! context is created lazily by `attach_provider`, which `GpuSteepestEdge::new` calls with the tableau's
//! columns before the first pivot; lazy providers (examples/max_flow.rs:149-163) are materialised there by
//! iterating `column(j)` once.  Columns passed to the query methods are recognised by content.
use std::cell::{Cell, RefCell};
use std::collections::HashMap;
use std::fmt;

use relp::algorithm::two_phase::matrix_provider::column::Column;
use relp::algorithm::two_phase::matrix_provider::filter::Filtered;
use relp::algorithm::two_phase::matrix_provider::MatrixProvider;
use relp::algorithm::two_phase::tableau::inverse_maintenance::{ops, ColumnComputationInfo, InverseMaintainer};
use relp::algorithm::two_phase::tableau::kind::Kind;
use relp::algorithm::two_phase::tableau::BasisChangeComputationInfo;
use relp::data::linear_algebra::traits::Element;
use relp::data::linear_algebra::vector::{DenseVector, SparseVector, Vector};
use relp::data::linear_algebra::SparseTuple;
use relp_num::RationalBig;

use crate::ffi;
use crate::number::{lcm, num_den, rational_from_limbs, sparse_from_limbs};

/// `ColumnComputationInfo`: the generated column stays on the device; the host copy is made on demand.
#[derive(Debug)]
pub struct GpuColumn {
    pub(crate) q: usize,
    column: SparseVector<RationalBig, RationalBig>,
}
impl ColumnComputationInfo<RationalBig> for GpuColumn {
    fn column(&self) -> &SparseVector<RationalBig, RationalBig> { &self.column }
    fn into_column(self) -> SparseVector<RationalBig, RationalBig> { self.column }
}

pub struct GpuCarry {
    ctx: Cell<*mut ffi::rg_context>,
    m: usize,
    /// integer image of the problem: row i was multiplied by `row_scale[i]`
    row_scale: RefCell<Vec<i128>>,
    rhs: Vec<(i128, i128)>,
    /// engine ids per row: provider column j -> j, artificial a -> a - nr_artificial
    basis_ids: RefCell<Vec<i32>>,
    nr_artificial: usize,
    column_ids: RefCell<HashMap<Vec<(usize, String)>, usize>>,
    b_mirror: RefCell<DenseVector<RationalBig>>,
}

impl GpuCarry {
    fn check(&self, rc: i32, what: &str) {
        if rc != ffi::RG_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(ffi::rg_last_error(self.ctx.get())) };
            panic!("{what} failed ({rc}): {}", msg.to_string_lossy());   // hot-path errors are panics in relp too
        }
    }
    pub(crate) fn ctx(&self) -> *mut ffi::rg_context {
        let c = self.ctx.get();
        assert!(!c.is_null(), "GpuCarry used before attach_provider (use GpuSteepestEdge / solve_relaxation_gpu)");
        c
    }
    fn limbs(&self) -> usize {
        let mut l = 0;
        self.check(unsafe { ffi::rg_get_limbs(self.ctx(), &mut l) }, "rg_get_limbs");
        l as usize
    }
    fn denominator(&self) -> Vec<u64> {
        let mut d = vec![0u64; self.limbs()];
        self.check(unsafe { ffi::rg_get_denominator(self.ctx(), d.as_mut_ptr()) }, "rg_get_denominator");
        d
    }
    fn refresh_b(&self) {
        let (l, d) = (self.limbs(), self.denominator());
        let mut w = vec![0u64; self.m * l];
        self.check(unsafe { ffi::rg_get_b(self.ctx(), w.as_mut_ptr()) }, "rg_get_b");
        let values = (0..self.m).map(|i| rational_from_limbs(&w[i * l..(i + 1) * l], &d)).collect();
        *self.b_mirror.borrow_mut() = DenseVector::new(values, self.m);
    }

    /// Uploads the provider's columns (integer prescale) and builds the identity carry.  Called once, with all
    /// provider columns in index order and the phase-one / phase-two cost of each (None: phase one).
    pub fn attach_provider<C: Column>(&self, columns: &[C], costs: Option<&[RationalBig]>)
    where C::F: fmt::Display {
        if !self.ctx.get().is_null() { return; }
        let m = self.m;
        // row scale r_i = lcm of the denominators in row i and of b_i
        let mut scale: Vec<i128> = self.rhs.iter().map(|(_, d)| *d).collect();
        for c in columns { for (i, v) in c.iter() { scale[i] = lcm(scale[i], num_den(v).1); } }
        let (mut colptr, mut rowidx, mut vals) = (vec![0i64], Vec::<i32>::new(), Vec::<i64>::new());
        let mut ids = self.column_ids.borrow_mut();
        for (j, c) in columns.iter().enumerate() {
            let mut key = Vec::new();
            for (i, v) in c.iter() {
                let (n, d) = num_den(v);
                rowidx.push(i as i32);
                vals.push(i64::try_from(n * (scale[i] / d)).expect("prescaled coefficient exceeds 63 bits"));
                key.push((i, v.to_string()));
            }
            colptr.push(rowidx.len() as i64);
            ids.entry(key).or_insert(j);
        }
        let rhs: Vec<i64> = self.rhs.iter().zip(&scale)
            .map(|((n, d), r)| i64::try_from(n * (r / d)).expect("prescaled rhs exceeds 63 bits")).collect();
        let opts = ffi::rg_options { device: 0, initial_limbs: 0, rank: 0, world: 1, dense_carry: 0, reserved: 0,
                                     nccl_unique_id: std::ptr::null() };
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { ffi::rg_create(&opts, &mut ctx) };
        self.ctx.set(ctx);
        self.check(rc, "rg_create");
        self.check(unsafe { ffi::rg_load_csc(ctx, m as i32, columns.len() as i32, colptr.as_ptr(), rowidx.as_ptr(),
                                             vals.as_ptr()) }, "rg_load_csc");
        self.check(unsafe { ffi::rg_set_rhs(ctx, rhs.as_ptr()) }, "rg_set_rhs");
        // (rational inputs additionally need rg_set_weights: relp_b200/frontend.py::prescale is the reference
        //  implementation of the weight vectors; omitted here for integer-row providers, all weights 1)
        let cost_int: Option<Vec<i64>> = costs.map(|cs| cs.iter().map(|c| {
            let (n, d) = num_den(c); assert_eq!(d, 1, "integer costs expected after the cost prescale"); n as i64
        }).collect());
        self.check(unsafe { ffi::rg_init_identity_basis(ctx, self.basis_ids.borrow().as_ptr(),
                                                        cost_int.as_ref().map_or(std::ptr::null(), |c| c.as_ptr())) },
                   "rg_init_identity_basis");
        *self.row_scale.borrow_mut() = scale;
        self.refresh_b();
    }

    fn column_index<C: Column>(&self, column: &C) -> usize where C::F: fmt::Display {
        let key: Vec<(usize, String)> = column.iter().map(|(i, v)| (i, v.to_string())).collect();
        *self.column_ids.borrow().get(&key).expect("column was not part of the attached provider")
    }

    fn unattached(m: usize, rhs: Vec<(i128, i128)>, basis_ids: Vec<i32>, nr_artificial: usize) -> Self {
        Self { ctx: Cell::new(std::ptr::null_mut()), m, row_scale: RefCell::new(vec![1; m]), rhs,
               basis_ids: RefCell::new(basis_ids), nr_artificial, column_ids: RefCell::new(HashMap::new()),
               b_mirror: RefCell::new(DenseVector::new(Vec::new(), 0)) }
    }
}

impl Drop for GpuCarry {
    fn drop(&mut self) { if !self.ctx.get().is_null() { unsafe { ffi::rg_destroy(self.ctx.get()); } } }
}
impl fmt::Display for GpuCarry {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result { write!(f, "GpuCarry(m = {})", self.m) }
}

impl InverseMaintainer for GpuCarry {
    type F = RationalBig;
    type ColumnComputationInfo = GpuColumn;

    // carry/mod.rs:374-395
    fn create_for_fully_artificial<Rhs: Element>(rhs: DenseVector<Rhs>) -> Self where Self::F: ops::Rhs<Rhs> {
        let m = rhs.len();
        let b = rhs.iter().map(|v| num_den(v)).collect();
        Self::unattached(m, b, (0..m).map(|a| a as i32 - m as i32).collect(), m)
    }
    // carry/mod.rs:397-442: basis_indices are in relp's phase-one index space (artificials first)
    fn create_for_partially_artificial<G: Element>(artificial: &[usize], _basis: &[(usize, usize)],
                                                   b: DenseVector<G>, basis_indices: Vec<usize>) -> Self
    where Self::F: ops::Rhs<G> {
        let na = artificial.len();
        let ids = basis_indices.iter().map(|&j| j as i32 - na as i32).collect();
        Self::unattached(b.len(), b.iter().map(|v| num_den(v)).collect(), ids, na)
    }
    // carry/mod.rs:444-478: a general basis is inverted on the device (fraction-free Gauss-Jordan, rg_init_basis)
    fn from_basis<'a, MP: MatrixProvider>(basis: &[usize], provider: &'a MP) -> Self {
        let m = provider.nr_rows();
        let carry = Self::unattached(m, provider.right_hand_side().iter().map(|v| num_den(v)).collect(),
                                     (0..m).map(|a| a as i32 - m as i32).collect(), 0);
        let columns: Vec<_> = (0..provider.nr_columns()).map(|j| provider.column(j)).collect();
        carry.attach_provider(&columns, None);
        let (ids, cost): (Vec<i32>, Vec<i64>) = (basis.iter().map(|&j| j as i32).collect(),
            (0..provider.nr_columns()).map(|j| num_den(&provider.cost_value(j)).0 as i64).collect());
        carry.check(unsafe { ffi::rg_init_basis(carry.ctx(), ids.as_ptr(), cost.as_ptr()) }, "rg_init_basis");
        *carry.basis_ids.borrow_mut() = ids;
        carry.refresh_b();
        carry
    }
    fn from_basis_pivots<'a, MP: MatrixProvider>(basis: &[(usize, usize)], provider: &'a MP) -> Self {
        let mut elements = basis.to_vec();
        elements.sort_by_key(|&(row, _)| row);
        Self::from_basis(&elements.into_iter().map(|(_, c)| c).collect::<Vec<_>>(), provider)
    }
    // carry/mod.rs:499-525: install the phase-two costs, rebuild -pi and -obj on the device
    fn from_artificial<'p, MP: MatrixProvider>(artificial: Self, provider: &'p MP, nr_artificial: usize) -> Self {
        debug_assert_eq!(nr_artificial, artificial.nr_artificial);
        let cost: Vec<i64> = (0..provider.nr_columns()).map(|j| num_den(&provider.cost_value(j)).0 as i64).collect();
        artificial.check(unsafe { ffi::rg_phase_switch(artificial.ctx(), cost.as_ptr()) }, "rg_phase_switch");
        artificial
    }
    // carry/mod.rs:527-559: redundant rows stay in the device carry as inert rows (value-identical, see
    // include/relp_gpu.h at rg_phase_switch); only the host index maps shrink
    fn from_artificial_remove_rows<'a, MP: Filtered>(artificial: Self, rows_removed: &'a MP, nr_artificial: usize) -> Self {
        Self::from_artificial(artificial, rows_removed, nr_artificial)
    }

    // carry/mod.rs:561-604
    fn change_basis<K: Kind>(&mut self, pivot_row_index: usize, pivot_column_index: usize, column: GpuColumn,
                             _cost: RationalBig, _kind: &K) -> BasisChangeComputationInfo<RationalBig> {
        let q = pivot_column_index as i32 - self.nr_artificial as i32;
        let mut info = ffi::rg_pivot_info::default();
        self.check(unsafe { ffi::rg_bring_into_basis(self.ctx(), q, pivot_row_index as i32, 1, &mut info) },
                   "rg_bring_into_basis");
        self.basis_ids.borrow_mut()[pivot_row_index] = q;
        self.refresh_b();
        // the three vectors of the info, for stock host-side pivot rules (the device rule keeps its own)
        let l = self.limbs();
        let (mut col, mut work, mut row, mut d0) =
            (vec![0u64; self.m * (l + 2)], vec![0u64; self.m * (2 * l + 5)], vec![0u64; self.m * l], vec![0u64; l]);
        self.check(unsafe { ffi::rg_get_basis_change_info(self.ctx(), col.as_mut_ptr(), work.as_mut_ptr(),
                                                          row.as_mut_ptr(), d0.as_mut_ptr()) }, "rg_get_basis_change_info");
        let d1 = self.denominator();
        let d0sq = square(&d0);
        let _ = column;
        BasisChangeComputationInfo {
            pivot_row_index, pivot_column_index,
            leaving_column_index: (info.leaving + self.nr_artificial as i32) as usize,
            column_before_change: SparseVector::new(sparse_from_limbs(&col, l + 2, self.m, &d0), self.m),
            work_vector: SparseVector::new(sparse_from_limbs(&work, 2 * l + 5, self.m, &d0sq), self.m),
            basis_inverse_row: SparseVector::new(sparse_from_limbs(&row, l, self.m, &d1), self.m),
        }
    }

    // carry/mod.rs:606-611: -pi . column = relative cost minus the column's own cost
    fn cost_difference<C: Column>(&self, original_column: &C) -> RationalBig where Self::F: ops::Column<C::F> {
        let (l, d) = (self.limbs(), self.denominator());
        let mut pi = vec![0u64; self.m * l];
        self.check(unsafe { ffi::rg_get_minus_pi(self.ctx(), pi.as_mut_ptr()) }, "rg_get_minus_pi");
        let scale = self.row_scale.borrow();
        original_column.iter().map(|(i, v)| {
            // -pi is exported for the row-scaled problem: (-pi_i r_i) (a_ij / 1); undo the scale per row
            let (n, dd) = num_den(v);
            rational_from_limbs(&pi[i * l..(i + 1) * l], &d) * RationalBig::from(scale[i] as i64) * ratio(n, dd)
        }).sum()
    }
    // carry/mod.rs:613-621
    fn generate_column<C: Column>(&self, original_column: C) -> GpuColumn where Self::F: ops::Column<C::F> {
        let q = self.column_index(&original_column);
        self.check(unsafe { ffi::rg_generate_column(self.ctx(), q as i32) }, "rg_generate_column");
        let (l, d) = (self.limbs(), self.denominator());
        let mut w = vec![0u64; self.m * (l + 2)];
        self.check(unsafe { ffi::rg_get_pivot_column(self.ctx(), w.as_mut_ptr()) }, "rg_get_pivot_column");
        GpuColumn { q, column: SparseVector::new(sparse_from_limbs(&w, l + 2, self.m, &d), self.m) }
    }
    // basis_inverse_rows.rs:179-195
    fn generate_element<C: Column>(&self, i: usize, original_column: C) -> Option<RationalBig> where Self::F: ops::Column<C::F> {
        let q = self.column_index(&original_column);
        let (l, d) = (self.limbs(), self.denominator());
        let mut w = vec![0u64; l + 2];
        self.check(unsafe { ffi::rg_get_element(self.ctx(), i as i32, q as i32, w.as_mut_ptr()) }, "rg_get_element");
        if w.iter().all(|x| *x == 0) { None } else { Some(rational_from_limbs(&w, &d)) }
    }
    // carry/mod.rs:636-645
    fn current_bfs(&self) -> Vec<SparseTuple<RationalBig>> {
        let b = self.b_mirror.borrow();
        let mut out: Vec<_> = self.basis_ids.borrow().iter().enumerate()
            .filter(|(i, _)| !num_traits::Zero::is_zero(&b[*i]))
            .map(|(i, id)| ((*id + self.nr_artificial as i32) as usize, b[i].clone())).collect();
        out.sort_by_key(|(j, _)| *j);
        out
    }
    fn basis_column_index_for_row(&self, row: usize) -> usize {
        (self.basis_ids.borrow()[row] + self.nr_artificial as i32) as usize
    }
    fn b(&self) -> DenseVector<RationalBig> { self.b_mirror.borrow().clone() }
    fn get_objective_function_value(&self) -> RationalBig {
        let (l, d) = (self.limbs(), self.denominator());
        let mut w = vec![0u64; l];
        self.check(unsafe { ffi::rg_get_minus_objective(self.ctx(), w.as_mut_ptr()) }, "rg_get_minus_objective");
        -rational_from_limbs(&w, &d)
    }
    fn get_constraint_value(&self, i: usize) -> &RationalBig {
        // SAFETY: the mirror is only replaced inside `&mut self` methods (change_basis) and the constructors
        unsafe { &(*self.b_mirror.as_ptr())[i] }
    }
}

fn ratio(n: i128, d: i128) -> RationalBig {
    use std::str::FromStr;
    RationalBig::from_str(&format!("{n}/{d}")).unwrap()
}
/// little-endian square of a positive limb array (the work vector lives over the squared denominator)
fn square(a: &[u64]) -> Vec<u64> {
    let mut r = vec![0u64; 2 * a.len()];
    for (i, &x) in a.iter().enumerate() {
        let mut carry = 0u128;
        for (j, &y) in a.iter().enumerate() {
            let t = x as u128 * y as u128 + r[i + j] as u128 + carry;
            r[i + j] = t as u64;
            carry = t >> 64;
        }
        r[i + a.len()] = carry as u64;
    }
    r
}
