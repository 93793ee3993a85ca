"""GPU parity: the CUDA engine (through the C ABI) against the CPU oracle, bit-exact.

Compared per case: status, the full pivot trace (phase, entering, row, leaving), iteration count,
exact objective and exact primal solution -- for all four pivot rules, through both the fused
`rg_iterate` loop and the trait-shaped call sequence.
"""
from fractions import Fraction as F

import numpy as np
import pytest

from oracle import relp_oracle as ro
from tests.common import oracle_trace, problem_from_provider, provider_from_problem

pytestmark = pytest.mark.gpu

RULE_NAMES = ["first_profitable", "first_profitable_with_memory", "dantzig", "steepest_edge"]


def check(provider=None, problem=None, rules=RULE_NAMES, modes=(True, False), initial_limbs=0):
    import relp_b200
    if problem is None:
        problem = problem_from_provider(provider)
    if provider is None:
        provider = provider_from_problem(problem)
    for rule in rules:
        ores, otrace = oracle_trace(provider, rule)
        for fused in modes:
            g = relp_b200.solve_relaxation(problem, rule=rule, fused=fused, initial_limbs=initial_limbs)
            tag = f"rule={rule} fused={fused}"
            assert g.status == ores.status, tag
            assert g.trace == otrace, tag
            assert g.pivots == len(otrace), tag
            if ores.status == "optimal":
                assert g.objective == ores.objective, tag
                assert g.bfs == ores.bfs, tag
            assert sorted(g.rows_removed) == sorted(ores.rows_removed), tag
    return g


def test_problem_2():
    from tests.test_oracle_golden import problem_2
    g = check(problem_2())
    assert g.objective == F(9, 2)
    assert g.bfs == [(1, F(1, 2)), (3, F(5, 2)), (4, F(3, 2))]


def test_problem_1():
    from tests.test_oracle_golden import problem_1
    g = check(problem_1())
    assert g.objective == 58
    assert g.bfs == [(0, F(4)), (2, F(6)), (5, F(2))]


def test_max_flow_example():
    adj = ro.adjacency_from_rows([[0, 0, 0, 0], [2, 0, 0, 0], [1, 1, 0, 0], [0, 1, 2, 0]])
    g = check(ro.MaxFlowPrimal(adj, 0, 3))
    dense = [0] * 10
    for j, v in g.bfs:
        dense[j] = v
    assert dense == [2, 1, 1, 1, 2, 0, 0, 0, 0, 0]


def test_shortest_path_example_fully_artificial():
    adj = ro.adjacency_from_rows([[0, 0, 0, 0], [1, 0, 0, 0], [2, 2, 0, 0], [0, 3, 1, 0]])
    g = check(ro.ShortestPathPrimal(adj, 0, 3))
    dense = [0] * 5
    for j, v in g.bfs:
        dense[j] = v
    assert dense == [0, 1, 0, 0, 1]


def random_matrix_data(rng, nv, counts, density=0.6, ub_prob=0.3, lo=-9, hi=9):
    """A random integer MatrixData (reference matrix_data.rs layout): eq, range, <=, >= rows."""
    n_eq, n_rng, n_up, n_lo = counts
    mc = n_eq + n_rng + n_up + n_lo
    rows = []
    for _ in range(mc):
        row = [int(rng.integers(lo, hi + 1)) if rng.random() < density else 0 for _ in range(nv)]
        rows.append(row)
    x0 = [int(rng.integers(0, 4)) for _ in range(nv)]       # a feasible point => feasible LP (mostly)
    b = []
    for i, row in enumerate(rows):
        ax = sum(a * x for a, x in zip(row, x0))
        if i < n_eq:
            v = ax
        elif i < n_eq + n_rng:
            v = ax + int(rng.integers(0, 3))
        elif i < n_eq + n_rng + n_up:
            v = ax + int(rng.integers(0, 5))
        else:
            v = ax - int(rng.integers(0, 5))
        if v < 0:                                           # make_b_non_negative
            rows[i] = [-a for a in row]
            v = -v
            # flipping an inequality swaps its type; keep the row in its group by re-drawing slack sign
            if i >= n_eq + n_rng:
                rows[i] = [-a for a in rows[i]]
                v = abs(ax) + 1 if i < n_eq + n_rng + n_up else max(0, abs(ax) - 1)
        b.append(v)
    ranges = [int(rng.integers(1, 6)) for _ in range(n_rng)]
    variables = [ro.Variable(int(rng.integers(-9, 10)),
                             int(rng.integers(1, 8)) if rng.random() < ub_prob else None)
                 for _ in range(nv)]
    cols = ro.columns_from_rows(rows, nv)
    return ro.MatrixData(cols, b, ranges, n_eq, n_rng, n_up, n_lo, variables)


@pytest.mark.parametrize("seed", range(12))
def test_random_small_lps(seed):
    rng = np.random.default_rng(1000 + seed)
    nv = int(rng.integers(3, 9))
    counts = tuple(int(rng.integers(0, 4)) for _ in range(4))
    if sum(counts) == 0:
        counts = (1, 0, 1, 0)
    check(random_matrix_data(rng, nv, counts))


@pytest.mark.parametrize("seed", range(4))
def test_random_rank_deficient(seed):
    rng = np.random.default_rng(2000 + seed)
    md = random_matrix_data(rng, 5, (3, 0, 1, 1), ub_prob=0.0)
    # duplicate an equality row => redundant constraint (Rank::Deficient path)
    cols = [[(i, v) for i, v in c] for c in md.constraint_columns]
    rows = [[0] * 5 for _ in range(len(md.b))]
    for j, c in enumerate(cols):
        for i, v in c:
            rows[i][j] = v
    rows.insert(1, list(rows[0]))
    b = list(md.b)
    b.insert(1, b[0])
    md2 = ro.MatrixData(ro.columns_from_rows(rows, 5), b, [], 4, 0, 1, 1, md.variables)
    check(md2)


@pytest.mark.parametrize("limbs", [1, 2, 4, 8, 10, 12, 14, 16])
def test_all_limb_widths_agree(limbs):
    rng = np.random.default_rng(77)
    check(random_matrix_data(rng, 7, (2, 1, 2, 1)), rules=["steepest_edge", "dantzig"], modes=(True,),
          initial_limbs=limbs)


@pytest.mark.parametrize("seed", range(3))
def test_bounded_lp_small_with_promotions(seed):
    """Synthetic recipe of configs 4/5 at oracle-sized dimensions; starts at 1 limb so the run
    crosses several width promotions (K9)."""
    from relp_b200.generators import bounded_lp
    import relp_b200
    prob = bounded_lp(40, 60, k_bounding=12, nnz_per_col=4, seed=seed)
    g = check(problem=prob, rules=["steepest_edge", "dantzig"], modes=(True,), initial_limbs=1)
    assert g.stats["promotions"] >= 1
    prob = bounded_lp(30, 40, k_bounding=10, dense=True, seed=seed)
    check(problem=prob, rules=["steepest_edge"], modes=(True,), initial_limbs=1)


def test_max_flow_random_graph():
    from relp_b200.generators import max_flow
    prob = max_flow(n_vertices=24, out_degree=3, seed=3, max_capacity=9)
    check(problem=prob, rules=["steepest_edge", "dantzig"], modes=(True,))


def test_row_sharded_two_gpus():
    """Row-sharded engine (NCCL) against the oracle on 2 GPUs of the box (skipped with fewer)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613",
                        os.path.join(root, "scripts", "mgpu_check.py")], capture_output=True, text=True,
                       timeout=900, cwd=root)
    assert r.returncode == 0 and "MGPU PARITY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_max_flow_300_against_fast_oracle():
    """Config 3 at a size the C++ oracle solves in a fraction of a second (m = 1494, phase one with
    ~300 artificial rows, zero-level pivots)."""
    import relp_b200
    from oracle import fast_oracle as fo
    from relp_b200.generators import max_flow
    prob = max_flow(300, 4, 0)
    ref = fo.solve_problem(prob, "steepest_edge")
    g = relp_b200.solve_relaxation(prob, rule="steepest_edge")
    assert g.status == ref.status == "optimal"
    assert g.trace == ref.trace
    assert g.objective == ref.objective and g.bfs == ref.bfs
    assert g.stats["limbs"] == 2 and g.denominator == 1      # totally unimodular: D stays 1


@pytest.mark.parametrize("seed", range(3))
def test_dense_int8_block_matches_csc_and_oracle(seed):
    """Config 5 data path at oracle size: structural columns as a dense int8 block (implicit indices)."""
    from relp_b200.generators import bounded_lp
    prob = bounded_lp(48, 64, k_bounding=14, dense=True, seed=seed, dense_block=True)
    assert prob.dense_block is not None
    for limbs in (1, 2):
        check(problem=prob, rules=["steepest_edge", "dantzig", "first_profitable"], modes=(True, False),
              initial_limbs=limbs)


@pytest.mark.parametrize("dense_carry", [0, 1])
def test_graph_replay_matches_eager_launches(dense_carry, monkeypatch):
    """The fused loop replays one CUDA graph per launch shape; RG_NO_GRAPH=1 enqueues the same kernels
    eagerly.  Both must give the oracle's trace (and so each other's), across promotions and list growth."""
    from relp_b200.generators import bounded_lp
    import relp_b200
    prob = bounded_lp(160, 240, k_bounding=16, nnz_per_col=4, seed=5)
    runs = []
    for no_graph in (False, True):
        if no_graph:
            monkeypatch.setenv("RG_NO_GRAPH", "1")
        else:
            monkeypatch.delenv("RG_NO_GRAPH", raising=False)
        for rule in ("steepest_edge", "dantzig"):
            g = relp_b200.solve_relaxation(prob, rule=rule, fused=True, initial_limbs=1, dense_carry=dense_carry)
            runs.append((rule, g.status, g.trace, g.objective, g.bfs))
    half = len(runs) // 2
    assert runs[:half] == runs[half:]
    for rule, status, trace, objective, bfs in runs[:half]:
        ores, otrace = oracle_trace(provider_from_problem(prob), rule)
        assert status == ores.status and trace == otrace and objective == ores.objective and bfs == ores.bfs


@pytest.mark.parametrize("limbs", [8, 12, 16])
def test_dense_block_tensor_core_dots_at_wide_limbs(limbs):
    """The dense dots run as byte-sliced u8 x s8 tensor-core products.  Starting at 8 / 16 limbs
    instantiates the widest slice-tile counts (sigma dot of 21 / 37 limbs: 22 / 38 tiles split over two CTA
    layers) on an oracle-sized problem."""
    from relp_b200.generators import bounded_lp
    prob = bounded_lp(48, 64, k_bounding=14, dense=True, seed=7, dense_block=True)
    check(problem=prob, rules=["steepest_edge"], modes=(True,), initial_limbs=limbs)


def test_dense_block_k_split_against_fast_oracle():
    """More than one 64-row chunk and more than one k-slice of the tensor-core dots (m = 200 -> 4 chunks),
    promotions from one limb, traces against the C++ oracle and against the CSC form of the same LP."""
    import relp_b200
    from oracle import fast_oracle as fo
    from relp_b200.generators import bounded_lp
    dense = bounded_lp(200, 120, k_bounding=20, dense=True, seed=3, dense_block=True)
    csc = bounded_lp(200, 120, k_bounding=20, dense=True, seed=3, dense_block=False)
    assert dense.dense_block is not None and csc.dense_block is None
    ref = fo.solve_problem(csc, "steepest_edge")
    for prob in (dense, csc):
        g = relp_b200.solve_relaxation(prob, rule="steepest_edge", initial_limbs=1)
        assert g.status == ref.status == "optimal"
        assert g.trace == ref.trace
        assert g.objective == ref.objective and g.bfs == ref.bfs
    assert g.stats["promotions"] >= 1


@pytest.mark.parametrize("start", [16, 10, 4])
def test_width_demotion_walks_down_and_up_again(start, monkeypatch):
    """The persistent state is narrowed when the numerators fit the next width with a margin (demotion) and widened
    again by the overflow prediction (K9): starting far too wide, with demotion allowed down to one limb, the run
    must demote at once (and promote again where the numbers grow past a width) and still walk the oracle's pivots exactly -- fused and
    trait-shaped, sparse and dense-block columns."""
    from relp_b200.generators import bounded_lp
    monkeypatch.setenv("RG_DEMOTE_FLOOR", "1")
    prob = bounded_lp(40, 60, k_bounding=12, nnz_per_col=4, seed=1)
    g = check(problem=prob, rules=["steepest_edge", "dantzig"], modes=(True, False), initial_limbs=start)
    assert g.stats["demotions"] >= 1, g.stats
    prob = bounded_lp(200, 120, k_bounding=20, dense=True, seed=3, dense_block=True)
    g = check(problem=prob, rules=["steepest_edge"], modes=(True,), initial_limbs=start)
    assert g.stats["demotions"] >= 1, g.stats
    rng = np.random.default_rng(5)
    check(random_matrix_data(rng, 9, (2, 2, 2, 1)), rules=["steepest_edge"], modes=(True, False), initial_limbs=start)
