"""The rectangular LMO of GWD-B (evrep_transport_plan_host, csrc/transport.cu) against scipy's LP solver - the path
oracle/gwd.py::_emd takes for n != m.  HOST code of the C-ABI library: runs without a GPU.

An optimal transportation plan need not be unique, so the comparison is on feasibility and on the objective value."""
import numpy as np
import pytest


def _lp(cost):
    from scipy.optimize import linprog
    n, m = cost.shape
    A = np.zeros((n + m, n * m))
    for i in range(n):
        A[i, i * m:(i + 1) * m] = 1
    for j in range(m):
        A[n + j, j::m] = 1
    b = np.r_[np.full(n, 1.0 / n), np.full(m, 1.0 / m)]
    res = linprog(cost.ravel().astype(np.float64), A_eq=A[:-1], b_eq=b[:-1], bounds=(0, None), method="highs")
    assert res.status == 0
    return res.fun


def _plan(cost):
    import event_representation_study_b200.batched as eb
    return eb.transport_plan_host(cost)


@pytest.mark.parametrize("n,m,kind", [(1, 1, "random"), (1, 7, "random"), (7, 1, "random"), (5, 8, "random"), (12, 18, "random"), (40, 25, "random"),
                                      (64, 63, "random"), (30, 30, "random"), (24, 36, "ties"), (17, 29, "constant"), (45, 60, "gw"),
                                      (100, 37, "gw")])
def test_transport_plan_is_feasible_and_optimal(n, m, kind):
    rng = np.random.default_rng(n * 131 + m)
    if kind == "random":
        cost = (rng.random((n, m)) * 10 - 3).astype(np.float32)
    elif kind == "ties":
        cost = rng.integers(0, 3, (n, m)).astype(np.float32)
    elif kind == "constant":
        cost = np.full((n, m), 1.5, np.float32)
    else:  # the structured cost of a conditional-gradient step: constC - hC1 G hC2^T at the product plan
        from oracle import gwd as ogwd
        Xs, Xt = rng.random((n, 4)), rng.random((m, 6))
        Ks, Kt = ogwd.compute_kernel(ogwd.pairwise_euclidean(Xs), ogwd.pairwise_euclidean(Xt), 0.7)
        p, q = np.ones(n) / n, np.ones(m) / m
        constC, hC1, hC2 = ogwd.gw_kl_init(Ks, Kt, p, q)
        cost = (constC - hC1 @ np.outer(p, q) @ hC2.T).astype(np.float32)
    G = _plan(cost)
    assert G.shape == (n, m) and (G >= 0).all()
    assert np.abs(G.sum(1) - 1.0 / n).max() < 1e-12 and np.abs(G.sum(0) - 1.0 / m).max() < 1e-12
    if kind != "ties":
        assert (G > 0).sum() <= n + m - 1  # a vertex of the transportation polytope
    got = float((cost.astype(np.float64) * G).sum())
    want = _lp(cost)
    assert abs(got - want) <= 1e-9 * max(1.0, np.abs(cost).max()), (got, want)


def test_transport_plan_larger_problem_runs_fast():
    import time
    rng = np.random.default_rng(0)
    cost = rng.random((700, 451)).astype(np.float32)
    t0 = time.perf_counter()
    G = _plan(cost)
    dt = time.perf_counter() - t0
    assert np.abs(G.sum(1) - 1.0 / 700).max() < 1e-12 and np.abs(G.sum(0) - 1.0 / 451).max() < 1e-12
    # dual feasibility check without an LP: the plan must beat 50 random feasible plans (products of marginals of random permutations)
    obj = float((cost.astype(np.float64) * G).sum())
    assert obj <= float(cost.mean()) + 1e-12  # the product plan p q^T is feasible
    assert dt < 20.0, dt


def test_transport_plan_rejects_non_finite_costs():
    from event_representation_study_b200._lib import EvrepError, EINVAL
    cost = np.ones((3, 4), np.float32)
    cost[1, 2] = np.nan
    with pytest.raises(EvrepError) as e:
        _plan(cost)
    assert e.value.code == EINVAL
