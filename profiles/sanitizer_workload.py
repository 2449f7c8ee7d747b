# small end-to-end exercise of every kernel family for compute-sanitizer
import numpy as np, torch, sys
sys.path.insert(0, '.')
import event_representation_study_b200.batched as eb
from event_representation_study_b200.synth import poisson_window
H, W = 120, 152
wins = [poisson_window(10 + i, n, H, W) for i, n in enumerate((30000, 17, 9001, 0, 20000))]
wins = [w for w in wins if len(w["x"])]
ev = eb.pack_events(wins, "cuda")
r = eb.ergo12(ev, H, W); eb.event_stack(ev, H, W, 12); eb.time_surface(ev, H, W, 6, 50000.0); eb.tore(ev, H, W, 6)
eb.event_stack(ev, H, W, 5); eb.tore(ev, H, W, 3)
eb.voxel_grid(ev, H, W, 5, "evlicious", normalize=True); eb.histogram(ev, H, W)
eb.detector_input(r, 160); eb.detector_input(r, 64, mode="squash")
A = torch.rand(200, 77, device="cuda"); B = torch.rand(130, 77, device="cuda"); eb.gemm_nt_3xtf32(A, B)
rng = np.random.default_rng(0); Xs = rng.random((40, 4)); Xt = rng.random((40, 6))
print(eb.gw_kl(Xs, Xt, 0.7, max_iter=5))
print(eb.gwd_kernel_l1([Xs], [Xt], 0.7))
torch.cuda.synchronize(); print("sanitizer workload done")
