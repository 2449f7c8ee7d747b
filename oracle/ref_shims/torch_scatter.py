"""Stand-in for the third-party `torch_scatter` package (absent offline; unpinned in the
reference, README.md:31) so that /root/reference/representations/representation_search/operations.py
can be imported and executed UNMODIFIED when generating golden vectors.

TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py in the build container).

Semantics restated from torch_scatter 2.1 (`scatter(src, index, dim=-1, dim_size, reduce)`):
  sum  : out[i] = sum of src[j] with index[j] == i, 0 for untouched i
  mean : sum / clamp(count, min=1)   (true division for floating src)
  max  : max of src[j]; untouched i -> 0
  min  : min of src[j]; untouched i -> 0
"""
import torch


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert src.dim() == 1 and index.dim() == 1, "shim covers the 1-D call sites only"
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    if reduce in ("sum", "add"):
        return torch.zeros(dim_size, dtype=src.dtype).scatter_add_(0, index, src)
    if reduce == "mean":
        s = torch.zeros(dim_size, dtype=src.dtype).scatter_add_(0, index, src)
        c = torch.zeros(dim_size, dtype=src.dtype).scatter_add_(0, index, torch.ones_like(src))
        c = c.clamp(min=1)
        if src.is_floating_point():
            return s / c
        return torch.div(s, c, rounding_mode="floor")
    if reduce == "max":
        o = torch.zeros(dim_size, dtype=src.dtype)
        if index.numel():
            o.scatter_reduce_(0, index, src, "amax", include_self=False)
        return o
    if reduce == "min":
        o = torch.zeros(dim_size, dtype=src.dtype)
        if index.numel():
            o.scatter_reduce_(0, index, src, "amin", include_self=False)
        return o
    raise ValueError(reduce)


def _arg_of(src, index, values, dim_size):
    """torch_scatter's arg output: position of the element that set each output (the CPU kernel updates on a STRICT
    comparison, so among equal values the first one stays), `src.numel()` for untouched outputs."""
    n = src.numel()
    arg = torch.full((dim_size,), n, dtype=torch.long)
    if n:
        cand = torch.where(src == values[index], torch.arange(n), torch.full((n,), n))
        arg.scatter_reduce_(0, index, cand, "amin", include_self=True)
    return arg


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    """torch_scatter.scatter_max -> (values, argmax); untouched entries: value 0, arg = src.numel()"""
    v = scatter(src, index, dim, out, dim_size, "max")
    return v, _arg_of(src, index, v, v.numel())


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    assert src.dim() == 1 and index.dim() == 1
    o = torch.zeros(dim_size, dtype=src.dtype)
    if index.numel():
        o.scatter_reduce_(0, index, src, "amin", include_self=False)
    return o, _arg_of(src, index, o, dim_size)
