# one LMO of a GW step at n = 1000 (the structured cost of profiles' GWD-B numbers), for ncu captures of k_auction
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import event_representation_study_b200.batched as eb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
rng = np.random.default_rng(56)
Xs = rng.random((n, 4))
Xt = np.concatenate([Xs[rng.permutation(n)][:, :3] + 0.05 * rng.standard_normal((n, 3)), rng.random((n, 11)) * 0.2], 1)
def kern(X):
    D2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    std = np.sqrt(D2.mean() / 2)
    return np.exp(-(np.sqrt(D2) / (0.7 * std)) ** 2 / 2)
Ks, Kt = kern(Xs), kern(Xt)
p = np.ones(n) / n
f1 = Ks * np.log(Ks + 1e-15) - Ks
constC = (f1 @ p)[:, None] + (Kt @ p)[None, :]
hC2 = np.log(Kt + 1e-15)
G = 0.5 * np.outer(p, p) + 0.5 * np.eye(n)[rng.permutation(n)] / n
cost = torch.as_tensor((constC - Ks @ G @ hC2.T).astype(np.float32)).cuda()
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sigma, st = eb.assignment_auction(cost)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(n, 'auction ms', round(dt * 1e3, 3), st, 'us/round', round(dt * 1e6 / max(st['rounds'], 1), 2))
