"""Writes tests/golden/est_*.npz with the reference's own ValueLayer / QuantizationLayer (learned_repr.py).  Runs only where
/root/reference exists.  The reference builds its MLP on "cuda" and returns `.cuda()`: both are patched to CPU no-ops.
TEST INFRASTRUCTURE ONLY (fixture generator; nothing in the product package imports it)."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import est as oest  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_learned_repr", "/root/reference/ev-YOLOv6/yolov6/models/learned_repr.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
torch.Tensor.cuda = lambda self, *a, **k: self          # forward() ends with .cuda()
torch.nn.Module.to = lambda self, *a, **k: self          # __init__ moves the value layer to "cuda"


def main():
    torch.manual_seed(0)
    for name, dim, S, counts in [("small", (6, 24, 30), 64, (3000, 1800)), ("tall_c3", (3, 40, 20), 48, (2500,))]:
        q = ref.QuantizationLayer(dim=dim, image_size=S)       # trains the value layer to the trilinear kernel (init_kernel)
        ws = [m.weight.detach().numpy().copy() for m in q.value_layer.mlp]
        bs = [m.bias.detach().numpy().copy() for m in q.value_layer.mlp]
        rng = np.random.default_rng(len(name))
        rows = []
        for b, n in enumerate(counts):
            t = np.sort(rng.integers(0, 90000, n)).astype(np.float32)
            rows.append(np.stack([rng.integers(0, dim[2], n), rng.integers(0, dim[1], n), t, rng.integers(0, 2, n), np.full(n, b)], 1))
        events = torch.tensor(np.concatenate(rows), dtype=torch.float32)
        with torch.no_grad():
            out = q.forward(events.clone())
            vox, lb = oest.est_forward(events.clone(), ws, bs, dim, S)
        assert torch.allclose(out, lb, rtol=1e-6, atol=1e-6), name             # the oracle restates the reference
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", f"est_{name}.npz"), events=events.numpy(), dim=np.array(dim), image_size=S,
                            out=out.numpy(), vox=vox.numpy(), **{f"w{i}": w for i, w in enumerate(ws)}, **{f"b{i}": b for i, b in enumerate(bs)})
        print(name, tuple(out.shape), float(out.abs().max()))


if __name__ == "__main__":
    main()
