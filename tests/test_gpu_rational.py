"""GPU parity on RATIONAL inputs: rows are prescaled to integers on the host and the engine
compensates with weights.  Traces, objectives and solutions must equal the
oracle's on the rational problem -- for Dantzig and steepest edge the weights are what makes that hold."""
import os
from fractions import Fraction as F

import numpy as np
import pytest

from oracle import fast_oracle as fo
from oracle import relp_oracle as ro
from tests.netlib_util import provider_from_mps, scaled_from_provider
from tests.test_oracle_golden import _two_var

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RULES = ["first_profitable", "first_profitable_with_memory", "dantzig", "steepest_edge"]


def check_rational(provider, rules=RULES, modes=(True, False), constant=F(0), expected=None):
    import relp_b200
    sp = scaled_from_provider(provider)
    for rule in rules:
        ref = fo.solve_provider(provider, rule)
        for fused, dense_carry in [(f, d) for f in modes for d in ((False, True) if f else (False,))]:
            g = relp_b200.solve_relaxation(sp.problem, rule=rule, fused=fused, dense_carry=dense_carry)
            tag = f"rule={rule} fused={fused} dense_carry={dense_carry}"
            assert g.status == ref.status, tag
            assert g.trace == ref.trace, tag
            if ref.status == "optimal":
                assert g.objective == ref.objective, tag
                assert g.bfs == ref.bfs, tag
                if expected is not None:
                    assert g.objective + constant == expected, tag
    return g


def test_two_phase_fixtures_with_fractions():
    # src/algorithm/two_phase/test.rs:46-212 (b = 3/2, 5/2; upper bound 3/4)
    cols = ro.columns_from_rows([[1, 0], [1, 1]], 2)
    data = ro.MatrixData(cols, [F(3, 2), F(5, 2)], [], 0, 0, 2, 0, [ro.Variable(-2), ro.Variable(-1)])
    g = check_rational(data)
    assert g.bfs == [(0, F(3, 2)), (1, F(1))]
    g = check_rational(_two_var([[1, 1], [1, 1], [1, 1]], [1, 1, 1], (3, 0, 0, 0)))
    assert g.bfs == [(0, F(3, 4)), (1, F(1, 4))]
    check_rational(_two_var([[1, 1], [0, 0]], [1, 0], (2, 0, 0, 0)))
    g = check_rational(_two_var([[1, 1], [0, 0]], [1, 1], (1, 0, 1, 0)))
    assert g.bfs == [(0, F(3, 4)), (1, F(1, 4)), (2, F(1))]


@pytest.mark.parametrize("seed", range(8))
def test_random_rational_lps(seed):
    from tests.test_gpu_parity import random_matrix_data
    rng = np.random.default_rng(4000 + seed)
    md = random_matrix_data(rng, int(rng.integers(3, 8)), (2, 1, 2, 1), ub_prob=0.4)
    dens = [1, 2, 3, 4, 5, 6, 10]
    # make the data rational: divide entries, right-hand sides, bounds and costs by small integers
    md.constraint_columns = [[(i, F(v, int(rng.choice(dens)))) for i, v in c] for c in md.constraint_columns]
    md.b = [F(b, int(rng.choice(dens))) for b in md.b]
    md.ranges = [F(r, int(rng.choice(dens))) for r in md.ranges]
    for v in md.variables:
        v.cost = F(v.cost, int(rng.choice(dens)))
        if v.upper_bound is not None:
            v.upper_bound = F(v.upper_bound, int(rng.choice(dens)))
    check_rational(md)


@pytest.mark.parametrize("name,expected", [
    ("afiro.mps", F(-406659, 875)),
    ("maros.mps", F(385, 3)),
    ("testprob.mps", F(54)),
])
def test_small_netlib(name, expected):
    lp, md = provider_from_mps(open(os.path.join(GOLD, name)).read())
    check_rational(md, constant=lp.constant, expected=expected)


def test_adlittle_exact():
    lp, md = provider_from_mps(open(os.path.join(GOLD, "adlittle.mps")).read())
    check_rational(md, rules=["steepest_edge", "dantzig"], modes=(True,), constant=lp.constant,
                   expected=F(24975305659811992079614961229, 120651674036153428931840))


def test_sc205_exact():
    lp, md = provider_from_mps(open(os.path.join(GOLD, "SC205.SIF")).read())
    g = check_rational(md, rules=["steepest_edge"], modes=(True,), constant=lp.constant)
    assert abs(float(g.objective + lp.constant) - (-5.220206121e+01)) < 1e-8
