"""POT `ot.gromov` entry points used by the reference, restated from the published POT 0.9 sources.
TEST INFRASTRUCTURE ONLY.

sampled_gromov_wasserstein (ot/gromov/_estimators.py in 0.9; ot/gromov.py in 0.8):
    T = outer(p, q); the `for cpt in range(max_iter)` mirror-descent loop is skipped when
    max_iter == 0; with log=True it then sets
        log['gw_dist_estimated'], log['gw_dist_std'] = GW_distance_estimation(C1, C2, p, q, loss_fun, T, ...)
    GW_distance_estimation samples index pairs and evaluates
        list_value_sample[:, :, n] = loss_fun(C1[np.ix_(index_k[:, n], index_k[:, n+? ])], C2[...])
    then returns mean and std over the samples.  The reference passes a `loss_fun` that IGNORES its
    arguments and returns abs(pad(Ks) - pad(Kt)) (compute_otmi.py:73-75), so the estimate is exactly
    mean(abs(pad(Ks) - pad(Kt))) regardless of what is sampled.  We therefore call loss_fun once on
    1x1 dummies and take the mean - the random sampling cannot influence the value.

gromov_wasserstein(C1, C2, p, q, 'kl_loss') (ot/gromov/_gw.py): conditional gradient, see oracle/gwd.py
    (`gw_kl_cg`) which this shim calls.
"""
import numpy as np


def sampled_gromov_wasserstein(C1, C2, p, q, loss_fun, nb_samples_grad=100, epsilon=1, max_iter=500,
                               log=False, verbose=False, random_state=None):
    assert max_iter == 0, "shim restates only the max_iter=0 path the reference uses"
    T = np.outer(p, q)
    if log:
        L = loss_fun(np.zeros((1, 1)), np.zeros((1, 1)))
        return T, {"gw_dist_estimated": float(np.mean(L)), "gw_dist_std": float(np.std(L))}
    return T


def gromov_wasserstein(C1, C2, p, q, loss_fun="square_loss", log=False, verbose=False, **kw):
    from oracle.gwd import gw_kl_cg
    assert loss_fun == "kl_loss"
    T, gw = gw_kl_cg(C1, C2, p, q)
    if log:
        return T, {"gw_dist": gw}
    return T
