"""CPU: the MPS front-end + oracles reproduce the exact optima the reference's integration tests hold
(see tests/golden/README.md), and the integer prescale preserves the problem."""
import os
from fractions import Fraction as F

import pytest

from oracle import fast_oracle as fo
from oracle import relp_oracle as ro
from tests.netlib_util import provider_from_mps, scaled_from_provider

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return provider_from_mps(open(os.path.join(GOLD, name)).read())


@pytest.mark.parametrize("name,expected", [
    ("afiro.mps", F(-406659, 875)),
    ("AFIRO.SIF", F(-406659, 875)),
    ("adlittle.mps", F(24975305659811992079614961229, 120651674036153428931840)),
    ("maros.mps", F(385, 3)),
    ("testprob.mps", F(54)),
])
def test_exact_optima_of_reference_tests(name, expected):
    lp, md = load(name)
    r = fo.solve_provider(md, "steepest_edge")
    assert r.status == "optimal"
    assert r.objective + lp.constant == expected


def test_afiro_python_oracle_agrees_with_fast_oracle():
    lp, md = load("afiro.mps")
    trace = ro.Trace()
    r = ro.solve_relaxation(md, "steepest_edge", trace)
    f = fo.solve_provider(md, "steepest_edge")
    assert r.objective == f.objective and r.bfs == f.bfs
    assert [(p, q, row, lv) for p, q, row, lv, _ in trace.pivots] == f.trace


def test_netlib_float_anchors():
    lp, md = load("ADLITTLE.SIF")
    r = fo.solve_provider(md, "steepest_edge")
    assert abs(float(r.objective + lp.constant) - 2.254949632e+05) < 1e-3     # tests/netlib/test.rs:24-28


def test_prescale_preserves_the_problem():
    for name in ("afiro.mps", "maros.mps", "adlittle.mps"):
        lp, md = load(name)
        sp = scaled_from_provider(md)
        r1 = fo.solve_provider(md, "dantzig")
        r2 = fo.solve_problem(sp.problem, "first_profitable")   # any rule: the optimum is unique
        assert r2.status == "optimal"
        assert r2.objective / sp.cost_scale == r1.objective
