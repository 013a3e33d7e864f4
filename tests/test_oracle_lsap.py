"""The C restatement of the assignment solver (oracle/lsap.c) against SciPy itself and the known answers of
SURVEY.md 8(c).  SciPy is the third-party dependency the reference calls (detr/matcher.py:94)."""
import numpy as np
import pytest
from scipy.optimize import linear_sum_assignment as scipy_lsa

from oracle import lsap_oracle as O


def _same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and b[0].dtype == np.int64


@pytest.mark.parametrize("kind", ["uniform", "ties", "zeros", "quarter"])
def test_random_sweeps_match_scipy(kind):
    rng = np.random.default_rng(0)
    for _ in range(150):
        nr, nc = int(rng.integers(1, 40)), int(rng.integers(1, 40))
        c = {"uniform": lambda: rng.random((nr, nc)), "ties": lambda: rng.integers(0, 3, (nr, nc)),
             "zeros": lambda: np.zeros((nr, nc)), "quarter": lambda: np.round(4 * rng.random((nr, nc))) / 4}[kind]()
        c = c.astype(np.float32)
        assert _same(scipy_lsa(c), O.linear_sum_assignment(c))
        assert _same(scipy_lsa(c.astype(np.float64)), O.linear_sum_assignment(c.astype(np.float64)))


def test_python_twin_matches_c():
    rng = np.random.default_rng(1)
    for _ in range(60):
        nr, nc = int(rng.integers(1, 20)), int(rng.integers(1, 20))
        c = rng.integers(0, 4, (nr, nc)).astype(np.float64)
        assert _same(O.linear_sum_assignment(c), O.linear_sum_assignment_py(c))


@pytest.mark.parametrize("m", [1, 10, 25, 50, 75, 100, 130])
def test_detr_shaped(m):
    rng = np.random.default_rng(m)
    for _ in range(3):
        c = (rng.random((100, m)) * 6 - 2).astype(np.float32)
        assert _same(scipy_lsa(c), O.linear_sum_assignment(c))


def test_known_answers():
    for shape in [(5, 3), (3, 5)]:
        r, c = O.linear_sum_assignment(np.zeros(shape))
        assert r.tolist() == [0, 1, 2] and c.tolist() == [0, 1, 2]
    r, c = O.linear_sum_assignment(np.zeros((100, 0)))
    assert r.size == 0 and c.size == 0
    with pytest.raises(ValueError, match="invalid numeric"):
        O.linear_sum_assignment(np.array([[np.nan, 1.0], [1.0, 2.0]]))
    with pytest.raises(ValueError, match="invalid numeric"):
        O.linear_sum_assignment(np.array([[-np.inf, 1.0], [1.0, 2.0]]))
    with pytest.raises(ValueError, match="infeasible"):
        O.linear_sum_assignment(np.array([[np.inf, np.inf], [1.0, 2.0]]))
    with pytest.raises(ValueError):
        scipy_lsa(np.array([[np.inf, np.inf], [1.0, 2.0]]))


def test_batch_entry_point():
    rng = np.random.default_rng(3)
    costs = [rng.random((100, m)).astype(np.float32) for m in (3, 0, 17, 100)]
    got = O.linear_sum_assignment_batch(costs)
    for c, g in zip(costs, got):
        assert _same(scipy_lsa(c), g)
