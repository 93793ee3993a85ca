"""relp_b200 -- B200-native exact simplex engine behind relp's MatrixProvider / PivotRule /
InverseMaintainer surface.  CUDA kernels + C ABI live in `csrc/`; this package is the thin
Python face used by the tests and the benchmark."""
from .solver import IntegerProblem, solve_relaxation, RULES  # noqa: F401
