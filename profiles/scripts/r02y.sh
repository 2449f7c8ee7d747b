cat > /tmp/wa.py <<'PY'
import torch, numpy as np, math
import event_representation_study_b200.batched as eb
dev = torch.device("cuda", 0)
B, S = 64, 640
lb = torch.rand((B, 12, S, S), device=dev) * 255
out = torch.empty_like(lb)
rng = np.random.default_rng(4)
Ms = np.tile(np.eye(3), (B, 1, 1))
for b in range(B):
    ang, sc = math.radians(rng.uniform(-10, 10)), rng.uniform(0.9, 1.1)
    R = np.array([[math.cos(ang) * sc, math.sin(ang) * sc, 0], [-math.sin(ang) * sc, math.cos(ang) * sc, 0], [0, 0, 1.0]])
    Cm, T = np.eye(3), np.eye(3); Cm[:2, 2] = -S / 2; T[:2, 2] = rng.uniform(0.4, 0.6, 2) * S
    Ms[b] = T @ R @ Cm
for _ in range(3): eb.augment_affine(lb, Ms, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eb.augment_affine(lb, Ms, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("warp ms", ms, "GB/s", 2 * lb.numel() * 4 / ms / 1e6)
PY
PYTHONPATH=$PWD python /tmp/wa.py
PYTHONPATH=$PWD ncu --set full --clock-control none -k regex:k_warp_affine -s 1 -c 1 python /tmp/wa.py 2>&1 | grep -E "Duration|DRAM Throughput|Issue Slots Busy|L1/TEX Hit|Registers Per|Achieved Occupancy|Stall|One or More|Local|L2 Cache Throughput|Mem Busy|Max Bandwidth" | head -20
timeout 600 python -m pytest tests/test_gpu_image.py -q 2>&1 | tail -2
