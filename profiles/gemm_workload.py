# one n = 4096 contraction through the TMA-fed path (for ncu)
import sys, torch
sys.path.insert(0, '.')
import event_representation_study_b200.batched as eb
n = 4096
A = torch.rand((n, n), device="cuda"); B = torch.rand((n, n), device="cuda"); out = torch.empty((n, n), device="cuda")
for _ in range(3):
    eb.gemm_nt_3xtf32(A, B, out=out)
torch.cuda.synchronize()
