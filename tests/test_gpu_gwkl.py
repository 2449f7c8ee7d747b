"""GWD-B on the GPU: the tcgen05 3xTF32 contraction against float64 numpy, and the conditional-gradient
Gromov-Wasserstein solve (gromov_wasserstein.py:39-69) against the numpy oracle (oracle/gwd.py::gw_kl_cg, the restated
POT algorithm; "POT parity unpinned").

Tolerances.  The contraction keeps 22 mantissa bits per operand (round to nearest) and drops the lo*lo term: every product
is within ~2^-22 of exact; k-blocks of 32 are summed in fp32 with round-to-nearest, so
|C - C64| <= 1e-6 * (|A| |B|^T) elementwise is the bar (also when every term has the same sign).  The GW loss is compared at 1e-5 relative (north_star's float
tolerance) on problems where both solvers take the same vertex sequence; the plan's marginals must be uniform to 1e-6
and the loss must not exceed the loss of the starting plan p q^T.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E(cuda_device):
    import event_representation_study_b200.batched as eb
    return eb


@pytest.mark.parametrize("packed", [True, False], ids=["tma-packed", "register-staged"])
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (256, 384, 96), (1000, 1000, 1000), (130, 70, 45), (1, 1, 1),
                                   (257, 129, 1001), (64, 300, 8)])
def test_gemm_nt_3xtf32_matches_float64(E, M, N, K, packed):
    import torch
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32)
    B = (rng.standard_normal((N, K)) * 3).astype(np.float32)
    rv = rng.standard_normal(M).astype(np.float32)
    cv = rng.standard_normal(N).astype(np.float32)
    got = E.gemm_nt_3xtf32(torch.as_tensor(A).cuda(), torch.as_tensor(B).cuda(), alpha=0.5, row_vec=torch.as_tensor(rv).cuda(),
                           col_vec=torch.as_tensor(cv).cuda(), packed=packed).cpu().numpy().astype(np.float64)
    A64, B64 = A.astype(np.float64), B.astype(np.float64)
    want = 0.5 * A64 @ B64.T + rv[:, None] + cv[None, :]
    bound = 1e-6 * (0.5 * np.abs(A64) @ np.abs(B64).T + np.abs(rv)[:, None] + np.abs(cv)[None, :]) + 1e-30
    err = np.abs(got - want)
    assert (err <= bound).all(), f"worst error / bound = {(err / bound).max():.2f} at {np.unravel_index((err / bound).argmax(), err.shape)}"


def test_gemm_kl_operand_ranges(E):
    """the operands of the GW tensor product: hC1 in (0, 1], hC2 = log(K + 1e-15) in [-34.5, 0], T ~ 1 / (n m)"""
    import torch
    rng = np.random.default_rng(5)
    n = 300
    hC1 = np.exp(-rng.random((n, n)) * 8).astype(np.float32)
    hC2 = np.log(np.exp(-rng.random((n, n)) * 40) + 1e-15).astype(np.float32)
    perm = rng.permutation(n)
    Bp = hC2[:, perm]
    got = E.gemm_nt_3xtf32(torch.as_tensor(hC1).cuda(), torch.as_tensor(Bp).cuda(), alpha=1.0 / n).cpu().numpy().astype(np.float64)
    want = hC1.astype(np.float64) @ Bp.astype(np.float64).T / n
    assert (np.abs(got - want) <= 1e-6 * np.abs(want)).all(), np.abs(got / want - 1).max()  # same-sign terms: |want| = sum |a||b| / n


def test_gemm_without_epilogue_vectors_and_unaligned_views(E):
    import torch
    rng = np.random.default_rng(9)
    A = torch.as_tensor(rng.standard_normal((200, 203)).astype(np.float32)).cuda()
    B = torch.as_tensor(rng.standard_normal((150, 203)).astype(np.float32)).cuda()
    got = E.gemm_nt_3xtf32(A, B).cpu().numpy().astype(np.float64)
    want = A.cpu().numpy().astype(np.float64) @ B.cpu().numpy().astype(np.float64).T
    assert np.abs(got - want).max() <= 1e-6 * (np.abs(A.cpu().numpy()).astype(np.float64) @ np.abs(B.cpu().numpy()).astype(np.float64).T).max()


def _clouds(seed, n, ds=4, dt=6):
    rng = np.random.default_rng(seed)
    Xs = rng.random((n, ds))
    # a noisy embedding of the same points plus distractor channels: a structured problem with a clear optimum
    Xt = np.concatenate([Xs[rng.permutation(n)][:, :3] + 0.05 * rng.standard_normal((n, 3)), rng.random((n, dt - 3))], 1)
    return Xs, Xt


def _loss64(Xs, Xt, P, h=0.7):
    """float64 GW-KL loss of a given plan and its Frank-Wolfe gap (exact LMO), from the oracle's building blocks"""
    from scipy.optimize import linear_sum_assignment
    from oracle import gwd as ogwd
    n = len(Xs)
    Ks, Kt = ogwd.compute_kernel(ogwd.pairwise_euclidean(Xs), ogwd.pairwise_euclidean(Xt), h)
    p = np.ones(n) / n
    constC, hC1, hC2 = ogwd.gw_kl_init(Ks, Kt, p, p)
    tens = constC - hC1 @ P @ hC2.T
    r, c = linear_sum_assignment(tens)
    Gc = np.zeros_like(P)
    Gc[r, c] = 1.0 / n
    return float(np.sum(tens * P)), float(2 * np.sum(tens * (P - Gc)))


@pytest.mark.parametrize("lmo", ["auction", "host"])
@pytest.mark.parametrize("n,seed", [(32, 1), (64, 2), (100, 3)])
def test_gw_kl_matches_oracle(E, n, seed, lmo):
    """problems on which the float32-operand solve takes the same vertex sequence as the float64 oracle"""
    from oracle import gwd as ogwd
    Xs, Xt = _clouds(seed, n)
    want = ogwd.gwd_b_cost(Xs, Xt, 0.7)
    got, iters, plan = E.gw_kl(Xs, Xt, 0.7, return_plan=True, lmo=lmo)
    assert iters >= 1
    P = plan.double().cpu().numpy()
    assert np.abs(P.sum(1) - 1.0 / n).max() < 1e-6 and np.abs(P.sum(0) - 1.0 / n).max() < 1e-6
    assert (P >= -1e-9).all()
    assert abs(got - want) <= 1e-5 * abs(want), (got, want, iters)


@pytest.mark.parametrize("n,seed", [(64, 2), (200, 4), (129, 5)])
def test_gw_kl_plan_is_a_stationary_point_with_the_reported_loss(E, n, seed):
    """Conditional gradient on this non-convex objective is chaotic in its rounding: one assignment that flips in the
    LMO (float32 operands vs the oracle's float64; n = 200 / seed 4 flips at the second step) leads to another local
    optimum, so the loss of the oracle's run is not a parity target in general.  What must hold for ANY run: the plan is in
    U(p, q), the reported loss is the float64 loss of that plan (1e-5), it is a Frank-Wolfe stationary point, and it is in
    the same range as the oracle's optimum."""
    from oracle import gwd as ogwd
    Xs, Xt = _clouds(seed, n)
    got, iters, plan = E.gw_kl(Xs, Xt, 0.7, return_plan=True)
    P = plan.double().cpu().numpy()
    assert np.abs(P.sum(1) - 1.0 / n).max() < 1e-6 and np.abs(P.sum(0) - 1.0 / n).max() < 1e-6 and (P >= -1e-9).all()
    loss64, gap = _loss64(Xs, Xt, P)
    assert abs(got - loss64) <= 1e-5 * abs(loss64), (got, loss64)
    assert gap <= 2e-5 * abs(loss64), gap  # no descent direction left (up to the float32 resolution of the gradient)
    want = ogwd.gwd_b_cost(Xs, Xt, 0.7)
    assert got <= 1.1 * want, (got, want)


def test_gw_kl_descends_from_the_product_plan(E):
    from oracle import gwd as ogwd
    n = 150
    Xs, Xt = _clouds(11, n)
    Ks, Kt = ogwd.compute_kernel(ogwd.pairwise_euclidean(Xs), ogwd.pairwise_euclidean(Xt), 0.7)
    p = np.ones(n) / n
    constC, hC1, hC2 = ogwd.gw_kl_init(Ks, Kt, p, p)
    G0 = np.outer(p, p)
    f0 = float(np.sum((constC - hC1 @ G0 @ hC2.T) * G0))
    got0, it0 = E.gw_kl(Xs, Xt, 0.7, max_iter=0)
    assert abs(got0 - f0) <= 1e-5 * abs(f0)  # zero iterations: the loss of p q^T, through the two-GEMM evaluation
    got, iters = E.gw_kl(Xs, Xt, 0.7)
    assert got <= f0 + 1e-9 and iters >= 1


@pytest.mark.parametrize("n,m,seed", [(10, 12, 0), (30, 21, 1), (64, 40, 2), (33, 100, 3), (96, 64, 4)])
def test_gw_kl_rectangular_matches_the_oracle(E, n, m, seed):
    """n != m (gromov_wasserstein.py:85-184 pairs N events with the non-empty pixels): the LMO is a transportation problem.
    The oracle solves it with scipy's linprog (oracle/gwd.py::_emd); equal losses where both take the same vertices,
    and in every case a feasible plan whose reported loss is its float64 loss and does not exceed the oracle's by > 10 %."""
    from oracle import gwd as ogwd
    rng = np.random.default_rng(seed)
    Xs = rng.random((n, 4))
    Xt = np.concatenate([rng.random((m, 5)) * 3, rng.random((m, 2))], 1)
    want = ogwd.gwd_b_cost(Xs, Xt, 0.7)
    st = {}
    got, iters, plan = E.gw_kl(Xs, Xt, 0.7, return_plan=True, stats=st)
    T = plan.double().cpu().numpy()
    assert T.shape == (n, m) and (T >= -1e-9).all()
    assert np.abs(T.sum(1) - 1.0 / n).max() <= 1e-6 / n and np.abs(T.sum(0) - 1.0 / m).max() <= 1e-6 / m
    Ks, Kt = ogwd.compute_kernel(ogwd.pairwise_euclidean(Xs), ogwd.pairwise_euclidean(Xt), 0.7)
    constC, hC1, hC2 = ogwd.gw_kl_init(Ks, Kt, np.ones(n) / n, np.ones(m) / m)
    f64 = float(np.sum((constC - hC1 @ T @ hC2.T) * T))
    assert abs(got - f64) <= 1e-5 * abs(f64), (got, f64)
    assert got <= want * 1.1 + 1e-12, (got, want, iters)
    assert st["host_fallbacks"] == iters or iters == 0  # every step's LMO ran on the host
    if n <= 30:
        assert abs(got - want) <= 1e-5 * abs(want), (got, want)


def test_otmi_mirror_rectangular(E):
    """the reference's own call shape: more events than pixels"""
    from event_representation_study_b200.representations.representation_search.gromov_wasserstein import OTMI
    rng = np.random.default_rng(8)
    Xs, Xt = rng.random((50, 4)), rng.random((35, 7))
    T, dist = OTMI(Xs, Xt, 0.7).solve()
    assert T.shape == (50, 35) and np.isfinite(dist)
    assert np.abs(T.sum(1) - 1 / 50).max() < 1e-7 and np.abs(T.sum(0) - 1 / 35).max() < 1e-7


def test_otmi_mirror_of_gromov_wasserstein_py(E):
    """reference call sequence (gromov_wasserstein.py:39-69): OTMI(Xs, Xt, h).solve() -> (T, gw_dist)"""
    from event_representation_study_b200.representations.representation_search.gromov_wasserstein import OTMI
    from oracle import gwd as ogwd
    Xs, Xt = _clouds(2, 64)
    T, dist = OTMI(Xs, Xt, 0.7).solve()
    assert T.shape == (64, 64) and T.dtype == np.float64
    assert abs(dist - ogwd.gwd_b_cost(Xs, Xt, 0.7)) <= 1e-5 * abs(dist)


def test_gw_kl_degenerate_input_is_nan_not_a_hang(E):
    """one point: std = 0, the reference's kernels are NaN (0 / 0); the solve must come back with NaN at once"""
    d, it = E.gw_kl(np.array([[0.1, 0.2, 0.3, 0.4]]), np.array([[0.5, 0.6]]), 0.7)
    assert np.isnan(d) and it == 0


@pytest.mark.parametrize("n,seed", [(2, 1), (7, 2), (64, 2), (129, 5)])
def test_auction_lmo_agrees_with_the_exact_host_lmo(E, n, seed):
    """same loss with the GPU auction as with the exact host assignment solver (on problems without a near-tie in
    the LMO; see the stationary-point test for why a flipped vertex is not an error), and no fallback to the host"""
    Xs, Xt = _clouds(seed, n, dt=6) if n >= 4 else (np.random.default_rng(seed).random((n, 4)), np.random.default_rng(seed + 1).random((n, 6)))
    st = {}
    da, ia = E.gw_kl(Xs, Xt, 0.7, lmo="auction", stats=st, max_iter=50)
    dh, ih = E.gw_kl(Xs, Xt, 0.7, lmo="host", max_iter=50)
    assert st["host_fallbacks"] == 0 and st["auction_rounds"] > 0
    assert abs(da - dh) <= 1e-6 * max(abs(dh), 1e-12), (da, dh, ia, ih, st)


@pytest.mark.parametrize("kind,n", [("random", 1), ("random", 2), ("random", 33), ("random", 500), ("random", 1500), ("gw", 300), ("gw", 1000),
                                    ("ties", 64), ("constant", 50)])
def test_assignment_auction_is_optimal(E, kind, n):
    """the device LMO against scipy's exact solver: a permutation whose cost is within n * eps_rel * range of the optimum"""
    import torch
    from scipy.optimize import linear_sum_assignment
    rng = np.random.default_rng(n)
    if kind == "random":
        cost = rng.random((n, n)).astype(np.float32) * 10 - 3
    elif kind == "ties":
        cost = rng.integers(0, 4, (n, n)).astype(np.float32)  # massive ties
    elif kind == "constant":
        cost = np.full((n, n), 2.5, np.float32)
    else:  # the structured matrix of a GW step: constC - hC1 G hC2^T at a mixed plan
        from oracle import gwd as ogwd
        Xs, Xt = _clouds(55, n, dt=14)
        Ks, Kt = ogwd.compute_kernel(ogwd.pairwise_euclidean(Xs), ogwd.pairwise_euclidean(Xt), 0.7)
        p = np.ones(n) / n
        constC, hC1, hC2 = ogwd.gw_kl_init(Ks, Kt, p, p)
        G = 0.5 * np.outer(p, p) + 0.5 * np.eye(n)[rng.permutation(n)] / n
        cost = (constC - hC1 @ G @ hC2.T).astype(np.float32)
    sigma, st = E.assignment_auction(torch.as_tensor(cost).cuda())
    sg = sigma.cpu().numpy()
    assert st["status"] == 0
    assert sorted(sg.tolist()) == list(range(n))
    r, c = linear_sum_assignment(cost.astype(np.float64))
    opt = cost.astype(np.float64)[r, c].sum()
    got = cost.astype(np.float64)[np.arange(n), sg].sum()
    rng_c = float(cost.max() - cost.min())
    assert got <= opt + n * 1e-9 * rng_c + 1e-12 * abs(opt), (got, opt, st)
