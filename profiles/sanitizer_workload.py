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
# round 2: fused order ops, sub-pixel voxel grid, packed wire format, rectangular GWD-B, EST forward / backward, filters, otmi preparation,
# the N-ImageNet rank loaders, a CUDA-graph replay (programmatic dependent launches inside a capture)
eb.order_ops_fused(ev, H, W)
eb.voxel_grid(ev, H, W, 5, "evlicious", normalize=True, divider=2)
from event_representation_study_b200 import packed, est
hx = {k: getattr(ev, k).cpu().numpy() for k in "xytp"}
pk = packed.pack_host(hx["x"].view(np.uint16), hx["y"].view(np.uint16), hx["t"], hx["p"], ev.offsets, H, W)
if pk is not None:
    eb.ergo12(packed.upload(pk), H, W)
print(eb.gw_kl(rng.random((30, 4)), rng.random((45, 6)), 0.7, max_iter=4))
ws_ = [rng.standard_normal((8, 1)), rng.standard_normal((8, 8)), rng.standard_normal((1, 8))]
bs_ = [rng.standard_normal(8), rng.standard_normal(8), rng.standard_normal(1)]
tabs = tuple(torch.as_tensor(v, dtype=torch.float64, device="cuda") for v in est.compile_value_layer(ws_, bs_))
tn = (ev.t.double() / float(ev.t.max())).float()
q = est.quantize(ev, H, W, 4, tabs, t_float=tn)
for kind, prm in (("refractory", 500.0), ("contrast", 2.0), ("resize", 0.0)):
    eb.filter_events(ev, H, W, kind, prm, fx=2 if kind == "resize" else 1, fy=2 if kind == "resize" else 1)
eo = np.stack([rng.integers(0, W, 4000), rng.integers(0, H, 4000), np.sort(rng.integers(0, 9000, 4000)), rng.choice([-1, 1], 4000)], 1).astype(np.int32)
eb.otmi_prepare(torch.tensor(eo), rng.random((64, 64, 3)) * (rng.random((64, 64, 1)) < 0.5), H, W, 64)
import event_representation_study_b200.n_imagenet as nimg
en = torch.tensor(np.stack([rng.integers(0, 32, 3000).astype(float), rng.integers(0, 24, 3000).astype(float), np.sort(rng.random(3000)) * 0.04 + 1.0, rng.choice([-1.0, 1.0], 3000)], 1))
nimg.reshape_then_acc_adj_sort(en, height=24, width=32)
# later round-2 additions: wire format 3, warp-affine augmentation, large INTER_AREA factors, GWD-A on packed point sets (band kernel)
pk3 = packed.pack_host(hx["x"].view(np.uint16), hx["y"].view(np.uint16), hx["t"], hx["p"], ev.offsets, H, W, fmt=3)
if pk3 is not None:
    eb.ergo12(packed.upload(pk3), H, W)
lb = eb.detector_input(r, 96, interp="linear", scale_out=1.0, reverse_channels=False)
Ms = np.tile(np.array([[1.02, 0.05, -3.0], [-0.04, 0.97, 4.0], [0, 0, 1.0]]), (ev.B, 1, 1))
eb.augment_affine(lb, Ms, [True] * ev.B, [False] * ev.B)
eb.detector_input(r, 16, mode="squash")
pa, pb = [rng.random((n_, 4)) for n_ in (130, 64, 1, 200)], [rng.random((n_, 6)) for n_ in (70, 64, 3, 333)]
(Xa, sa), (Xb, sb) = eb.gwd_pack(pa), eb.gwd_pack(pb)
print(eb.gwd_kernel_l1(Xa, Xb, 0.7, s_offsets=sa, t_offsets=sb))
out_g = torch.empty((ev.B, H, W, 12), device="cuda")
call = eb.GraphedCall(lambda: eb.ergo12(ev, H, W, out=out_g))
call.replay(); call.replay()
# run-time specialised mixed-density kernels (NVRTC): SBN with 12 channels, SBT with 9 (scalar staging), a fat plan on 512-pixel tiles,
# the interpreted kernel for comparison, and a hot tile through the specialised wide plan
wi, fu, ag = [2, 1, 3, 5, 0, 0, 6, 4, 0, 2, 4, 1], ["timestamp", "polarity", "count", "timestamp_pos", "timestamp_neg", "count_pos", "count_neg", "polarity",
                                                  "timestamp", "count", "timestamp_neg", "count_pos"], ["variance", "mean", "sum", "max", "mean", "mean", "sum", "variance", "sum", "max", "min", "mean"]
eb.mixed_density(ev, H, W, wi, fu, ag, "SBN", specialize=False)
eb.mixed_density(ev, H, W, wi, fu, ag, "SBN", specialize=True)
eb.mixed_density(ev, H, W, wi[:9], fu[:9], ag[:9], "SBT", specialize=True)
eb.mixed_density(ev, H, W, [0, 1, 2, 3, 4, 5, 6, 0, 0, 1, 2, 3], ["timestamp_pos"] * 7 + ["timestamp_neg", "count", "polarity", "count_pos", "timestamp"],
                 ["variance"] * 8 + ["sum", "mean", "sum", "max"], "SBN", specialize=True)
hot = poisson_window(99, 70000, 32, 32)
hot["x"][:] = 3; hot["y"][:] = 5
eb.mixed_density(eb.pack_events([hot], "cuda"), 32, 32, wi, fu, ag, "SBN", specialize=True)
torch.cuda.synchronize(); print("sanitizer workload done")
