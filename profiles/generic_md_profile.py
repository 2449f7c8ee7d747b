"""One interpreted mixed-density call on the headline batch (ncu target)."""
import torch
import event_representation_study_b200.batched as eb
from event_representation_study_b200.synth import device_batch
H, W, B, N = 720, 1280, 32, 1_000_000
dev = torch.device("cuda", 0)
d = device_batch(B, N, H, W, dev, seed=3)
ev = eb.EventBatch(d["x"], d["y"], d["t"], d["p"], d["offsets"].cpu().numpy())
wi = [0, 3, 2, 6, 5, 6, 2, 5, 1, 0, 4, 2]
fu = ["polarity", "timestamp_neg", "count_neg", "polarity", "count_pos", "count", "timestamp_pos", "count_neg", "timestamp_neg", "timestamp_pos", "timestamp", "count"]
ag = ["variance", "variance", "mean", "sum", "mean", "sum", "mean", "mean", "max", "max", "max", "mean"]
out = torch.empty((B, H, W, 12), device=dev)
for _ in range(3):
    eb.mixed_density(ev, H, W, wi, fu, ag, "SBN", out=out, specialize=False)
torch.cuda.synchronize()
