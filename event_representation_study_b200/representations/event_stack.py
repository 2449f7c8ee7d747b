"""Mirror of representations/event_stack.py (reference :5-137)."""
import numpy as np
import torch

from .. import batched as eb
from .._single import device, one_window


class _PreStacked(list):
    """What pre_stack returns: behaves like the reference's list of {'stacked_polarity': [...], 'index': [...]} dicts
    (materialised lazily from the dense GPU result), and lets post_stack reuse the dense tensors directly."""

    def __init__(self, dense, owner):
        self.dense = dense  # list of CUDA tensors (H, W, stack) float32, past [, future]
        self._owner = owner
        super().__init__([None] * len(dense))
        self._ready = False

    def _materialise(self):
        if not self._ready:
            for i, d in enumerate(self.dense):
                list.__setitem__(self, i, self._owner._sparse_from_dense(d))
            self._ready = True

    def __getitem__(self, i):
        self._materialise()
        return list.__getitem__(self, i)

    def __iter__(self):
        self._materialise()
        return list.__iter__(self)


class EventStack(object):
    NO_VALUE = 0.0
    STACK_LIST = ["stacked_polarity", "index"]

    def __init__(self, stack_size, num_of_event, height, width):
        self.stack_size = stack_size
        self.num_of_event = num_of_event
        self.height = height
        self.width = width

    # -- GPU core: dense (H, W, stack) of "sign of the latest event inside nested suffix window k" ------------
    def _dense(self, x, y, p_pm1, t):
        if len(x) == 0:
            raise ValueError("zero-size array to reduction operation minimum which has no identity")  # t.min() in the reference
        p = np.asarray(p_pm1)
        if p.size and not np.isin(p, (-1, 1)).all():
            raise ValueError("EventStack expects polarities in {0, 1} (mapped to -1/+1 by pre_stack)")
        ev = one_window(x, y, t, p, self.height, self.width)
        return eb.event_stack(ev, self.height, self.width, self.stack_size)[0]

    def _sparse_from_dense(self, dense):
        """The reference's sparse encoding (event_stack.py:84-114) of a dense stack."""
        d = dense.to(torch.int8)
        nxt = torch.cat([d[:, :, 1:], torch.zeros_like(d[:, :, :1])], dim=2)
        diff = (d - nxt).reshape(-1, self.stack_size)
        out = {"stacked_polarity": [], "index": []}
        for k in range(self.stack_size):
            idx = torch.nonzero(diff[:, k], as_tuple=False)[:, 0]
            out["index"].append(idx.cpu().numpy().astype(np.int32))
            out["stacked_polarity"].append(diff[idx, k].cpu().numpy())
        return out

    def pre_stack(self, event_sequence, last_timestamp):
        x = event_sequence["x"].astype(np.int32)
        y = event_sequence["y"].astype(np.int32)
        p = 2 * event_sequence["p"].astype(np.int8) - 1
        t = event_sequence["t"].astype(np.int64)
        assert len(x) == len(y) == len(p) == len(t)
        past = t <= last_timestamp
        dense = [self._dense(x[past], y[past], p[past], t[past])]
        future = t > last_timestamp
        if np.sum(future) != 0:  # reversed stream with flipped polarity (event_stack.py:28-39)
            dense.append(self._dense(x[future][::-1], y[future][::-1], -p[future][::-1], t[future][::-1]))
        return _PreStacked(dense, self)

    def post_stack(self, pre_stacked_event):
        if isinstance(pre_stacked_event, _PreStacked):
            dense = pre_stacked_event.dense
        else:  # a reference-style list of sparse dicts: replay the cumulative put on the GPU
            dense = []
            for pf in pre_stacked_event:
                cur = torch.zeros(self.height * self.width, dtype=torch.float32, device=device())
                planes = [None] * self.stack_size
                for k in range(self.stack_size - 1, -1, -1):
                    idx = torch.as_tensor(np.asarray(pf["index"][k]), device=cur.device).long()
                    val = torch.as_tensor(np.asarray(pf["stacked_polarity"][k]), device=cur.device).float()
                    cur[idx] = val
                    planes[k] = cur.clone()
                dense.append(torch.stack(planes, dim=1).reshape(self.height, self.width, self.stack_size))
        parts = [dense[0]]
        if len(dense) == 2:
            parts.append(torch.flip(dense[1], dims=[2]))
        return torch.stack(parts, dim=2).cpu().numpy()  # (H, W, 1 or 2, stack) float32

    def make_stack(self, x, y, p, t):
        return self._sparse_from_dense(self._dense(x, y, p, t))

    def stack_data(self, x, y, p, t_s):
        assert len(x) == len(y) == len(p) == len(t_s)
        if len(x) == 0:
            return {"stacked_polarity": np.zeros([self.height, self.width], dtype=np.int8)}
        ev = one_window(x, y, np.arange(len(x), dtype=np.int64), p, self.height, self.width)
        d = eb.event_stack(ev, self.height, self.width, 1)[0, :, :, 0]
        return {"stacked_polarity": d.to(torch.int8).cpu().numpy()}

    @staticmethod
    def collate_fn(batch):
        return torch.utils.data._utils.collate.default_collate(batch)
