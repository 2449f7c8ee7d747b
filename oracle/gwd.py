"""CPU ORACLE (numpy) for the Gromov-Wasserstein ranking distance ("GWD") path.

TEST INFRASTRUCTURE ONLY - see oracle/representations.py for the rules and the pinning statement.

GWD-A is what the paper pipeline calls (representation_search/compute_otmi.py): because the
reference passes max_iter=0 and a loss function that ignores its arguments, POT's
sampled_gromov_wasserstein(..., log=True)['gw_dist_estimated'] collapses to the closed form
    mean(abs(pad(Ks) - pad(Kt)))
(SURVEY.md 8a row a17).  POT is absent offline => "POT parity unpinned"; the closed form is pinned
against compute_otmi.py executed unmodified with the ot shim (oracle/ref_shims/ot) in gen_golden.py.

GWD-B (representation_search/gromov_wasserstein.py:62-69) is POT's conditional-gradient GW with
kl_loss, restated in gw_kl_cg below from the published POT 0.9 algorithm; LP vertex choice and line
search details are POT-version dependent => "POT parity unpinned".
"""
import numpy as np


def pairwise_euclidean(X):
    """sklearn.metrics.pairwise_distances(X, X) (compute_otmi.py:68-69), restated as the direct
    formula in float64 (sklearn's ||a||^2+||b||^2-2ab expansion differs at the 1e-8 level)."""
    X = np.asarray(X, np.float64)
    n, d = X.shape
    D2 = np.zeros((n, n))
    for k in range(d):
        diff = X[:, k][:, None] - X[:, k][None, :]
        D2 += diff * diff
    return np.sqrt(D2)


def compute_kernel(Cx, Cy, h):
    """compute_otmi.py:6-32 (== gromov_wasserstein.py:10-36)."""
    std1 = np.sqrt((Cx ** 2).mean() / 2)
    std2 = np.sqrt((Cy ** 2).mean() / 2)
    h1 = h * std1
    h2 = h * std2
    with np.errstate(divide="ignore", invalid="ignore"):
        Kx = np.exp(-((Cx / h1) ** 2) / 2)
        Ky = np.exp(-((Cy / h2) ** 2) / 2)
    return Kx, Ky


def gwd_a_cost(Xs, Xt, h=0.7):
    """OTMI(Xs, Xt, h).solve()[1] of compute_otmi.py:50-93 in closed form:
    (1/L^2) * sum_{i,j<L} |Ks_pad[i,j] - Kt_pad[i,j]|, L = max(n, m), zero padding bottom/right."""
    Ks, Kt = compute_kernel(pairwise_euclidean(Xs), pairwise_euclidean(Xt), h)
    n, m = Ks.shape[0], Kt.shape[0]
    L = max(n, m)
    A = np.zeros((L, L))
    B = np.zeros((L, L))
    A[:n, :n] = Ks
    B[:m, :m] = Kt
    return float(np.abs(A - B).mean())


def otmi_pairs(events, rep, height, width, rep_size):
    """The data preparation of otmi() (compute_otmi.py:96-203): returns the list of (Xs, Xt) pairs
    (three quadrants, densest dropped).  `events` is an integer (N,4) array [x,y,t,p]; like the torch
    int32 tensor the reference receives (gen1_compute.py:56-58) divisions are done in float32."""
    ev = np.asarray(events)
    X, Y = ev[:, 0], ev[:, 1]
    w2, h2 = width / 2 - 1, height / 2 - 1
    quads = [
        ev[(X >= 0) & (X <= w2) & (Y >= 0) & (Y <= h2)].copy(),
        ev[(X > w2) & (X <= width - 1) & (Y >= 0) & (Y <= h2)].copy(),
        ev[(X >= 0) & (X <= w2) & (Y > h2) & (Y <= height - 1)].copy(),
        ev[(X > w2) & (X <= width - 1) & (Y > h2) & (Y <= height - 1)].copy(),
    ]
    sizes = [q.shape[0] for q in quads]
    ind = sizes.index(max(sizes))
    for q in quads[1:]:
        q[:, 0] = q[:, 0] - q[:, 0].min()  # ValueError on an empty quadrant, like the reference
        q[:, 1] = q[:, 1] - q[:, 1].min()
    r = rep_size
    xys = [
        ([0, r // 2 - 1], [0, r / 2 - 1]),
        ([r / 2 - 1, r - 1], [0, r / 2 - 1]),
        ([0, r / 2 - 1], [r / 2 - 1, r - 1]),
        ([r / 2 - 1, r - 1], [r / 2 - 1, r - 1]),
    ]
    pairs = []
    f32 = np.float32
    for i, q in enumerate(quads):
        if i == ind:
            continue
        with np.errstate(divide="ignore", invalid="ignore"):
            x = q[:, 0].astype(f32) / f32((width - 1) // 2)
            y = q[:, 1].astype(f32) / f32((height - 1) // 2)
            t = q[:, 2]
            t = (t - t[0]).astype(f32) / f32(t[-1] - t[0])
            p = q[:, 3]
            p = (p - p.min()).astype(f32) / f32(p.max() - p.min())
        mask = (q[:, 0] < (width - 1) // 2) & (q[:, 1] < (height - 1) // 2)
        Xs = np.stack([x[mask], y[mask], t[mask], p[mask]], axis=-1)
        cx, cy = xys[i]
        rp = rep[int(cy[0]): int(cy[1]) + 1, int(cx[0]): int(cx[1]) + 1, :]
        xe = np.repeat(np.arange(0, rp.shape[0]).reshape(rp.shape[0], 1), rp.shape[1], axis=1) / (rp.shape[0] - 1)
        ye = np.repeat(np.arange(0, rp.shape[1]).reshape(1, rp.shape[1]), rp.shape[0], axis=0) / (rp.shape[1] - 1)
        rp = np.concatenate((rp, xe[..., None], ye[..., None]), axis=2).reshape((-1, rep.shape[2] + 2))
        rp = rp[np.abs(rp[:, :-2]).sum(-1) > 0]
        pairs.append((Xs.copy(), rp.copy()))
    return pairs


def otmi(events, rep, height, width, rep_size, h=0.7):
    """otmi() (compute_otmi.py:96-211): mean GWD-A cost over the three kept quadrants."""
    return float(np.mean([gwd_a_cost(Xs, Xt, h) for Xs, Xt in otmi_pairs(events, rep, height, width, rep_size)]))


# ----------------------------------------------------------------------------------------------
# GWD-B: conditional-gradient Gromov-Wasserstein with KL loss (restated POT algorithm)
# ----------------------------------------------------------------------------------------------
def _emd(a, b, M):
    """Exact OT plan for cost M.  n == m with uniform marginals -> assignment problem."""
    from scipy.optimize import linear_sum_assignment, linprog
    n, m = M.shape
    if n == m and np.allclose(a, a[0]) and np.allclose(b, b[0]):
        r, c = linear_sum_assignment(M)
        G = np.zeros((n, m))
        G[r, c] = a[0]
        return G
    A_eq = np.zeros((n + m, n * m))
    for i in range(n):
        A_eq[i, i * m:(i + 1) * m] = 1
    for j in range(m):
        A_eq[n + j, j::m] = 1
    res = linprog(M.ravel(), A_eq=A_eq[:-1], b_eq=np.r_[a, b][:-1], bounds=(0, None), method="highs")
    return res.x.reshape(n, m)


def gw_kl_init(C1, C2, p, q):
    """ot.gromov.init_matrix(..., 'kl_loss'): f1(a)=a log(a+1e-15)-a, f2(b)=b, h1(a)=a, h2(b)=log(b+1e-15)."""
    f1 = C1 * np.log(C1 + 1e-15) - C1
    constC = (f1 @ p)[:, None] + (C2 @ q)[None, :]
    return constC, C1, np.log(C2 + 1e-15)


def gw_kl_cg(C1, C2, p, q, max_iter=10000, tol_rel=1e-9, tol_abs=1e-9):
    """min_T <constC - hC1 T hC2^T, T> by conditional gradient with exact LMO and exact quadratic
    line search; returns (T, gw_dist)."""
    constC, hC1, hC2 = gw_kl_init(C1, C2, p, q)
    G = np.outer(p, q)

    def tens(T):
        return constC - hC1 @ T @ hC2.T

    f_val = float(np.sum(tens(G) * G))
    for _ in range(int(max_iter)):
        old = f_val
        tG = tens(G)
        Mi = 2 * tG
        Mi = Mi + Mi.min()
        Gc = _emd(p, q, Mi)
        dG = Gc - G
        dot = hC1 @ dG @ hC2.T
        a = -float(np.sum(dot * dG))
        b = float(np.sum(constC * dG)) - float(np.sum((hC1 @ G @ hC2.T) * dG)) - float(np.sum(dot * G))
        if a > 0:
            alpha = min(1.0, max(0.0, -b / (2 * a)))
        else:
            alpha = 1.0 if a + b < 0 else 0.0
        G = G + alpha * dG
        f_val = old + a * alpha ** 2 + b * alpha
        d = abs(f_val - old)
        if d < tol_abs or d / max(abs(f_val), 1e-300) < tol_rel:
            break
    return G, float(np.sum(tens(G) * G))


def gwd_b_cost(Xs, Xt, h=0.7):
    """OTMI(Xs, Xt, h).solve()[1] of gromov_wasserstein.py:39-69."""
    Ks, Kt = compute_kernel(pairwise_euclidean(Xs), pairwise_euclidean(Xt), h)
    n, m = len(Ks), len(Kt)
    return gw_kl_cg(Ks, Kt, np.ones(n) / n, np.ones(m) / m)[1]
