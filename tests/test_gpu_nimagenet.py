"""The N-ImageNet loader wrappers (n_imagenet/real_cnn_model/data/imagenet.py:1002-1134, SURVEY.md 8b caller ii; kept as
caller code in tests/nimagenet_callers.py: the reference's call sequence over this package's mirrors) and the upstream
count / time representations of the drop-in module, against outputs of the reference's own wrappers (tests/golden/nimg_wrappers.npz, made by
oracle/gen_golden_nimagenet.py; the time-surface wrapper of the reference cannot run - see that script - and is pinned
through the reference's ToTimesurface with the casts the wrapper intends)."""
import numpy as np
import pytest

from conftest import assert_close, golden, load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def N(cuda_device):
    import event_representation_study_b200.n_imagenet as n
    return n


@pytest.fixture(scope="module")
def C(cuda_device):
    import nimagenet_callers
    return nimagenet_callers


@pytest.fixture(scope="module")
def G():
    return load(golden("nimg_wrappers")[0][1])


def test_fix_events_training(C, G):
    s = C.fix_events_training(G["events_s"].copy())
    assert s.dtype.names == ("x", "y", "t", "p") and all(s.dtype[k] == np.dtype("<f8") for k in s.dtype.names)
    assert np.array_equal(s["t"], G["events_s"][:, 2])


@pytest.mark.parametrize("name,src,rtol,atol", [
    ("voxel_grid", "events_s", 1e-5, 2e-6),
    ("optimized", "events_us", 1e-5, 2e-7),
    ("event_stack", "events_s", 0, 0),
    ("tore", "events_us", 1e-5, 1e-6),
    ("time_surface", "events_us", 1e-5, 1e-30),
    ("to_image", "events_s", 0, 0),
])
def test_reshape_then_wrappers_match_the_reference(C, G, name, src, rtol, atol):
    import torch
    H, W = int(G["H"]), int(G["W"])
    rep = getattr(C, "reshape_then_" + name)(torch.tensor(G[src].copy()), height=H, width=W)
    assert torch.is_tensor(rep) and rep.dtype == torch.float32 and not rep.is_cuda
    want = G[name]
    assert tuple(rep.shape) == want.shape
    if rtol == 0:
        assert np.array_equal(rep.numpy(), want)
    else:
        assert_close(rep.numpy(), want, rtol=rtol, atol=atol, what=name)


def test_augment_hook_is_applied_first(N, G):
    import torch
    H, W = int(G["H"]), int(G["W"])
    flip = lambda e: torch.stack([W - 1 - e[:, 0], e[:, 1], e[:, 2], e[:, 3]], 1)  # noqa: E731
    a = N.reshape_then_acc_count_pol(torch.tensor(G["events_s"].copy()), augment=flip, height=H, width=W)
    b = N.reshape_then_acc_count_pol(torch.tensor(G["events_s"].copy()), height=H, width=W)
    assert np.array_equal(a.numpy(), b.numpy()[:, :, ::-1])


def test_count_loaders_accept_degenerate_samples(N):
    """imagenet.py:296-343 never read the time column (count_only not even the polarity): a one-event, zero-span, unsorted or
    empty sample gives a valid count image like torch.bincount(minlength=H * W) does"""
    import torch
    one = torch.tensor([[2.0, 1.0, 0.5, 1.0]], dtype=torch.float64)
    r = N.reshape_then_acc_count_only(one, height=4, width=4)
    assert float(r.sum()) == 1.0 and float(r[0, 1, 2]) == 1.0
    flat_t = torch.tensor([[0.0, 0.0, 0.3, 1.0], [1.0, 0.0, 0.3, -1.0], [1.0, 0.0, 0.1, 0.0]], dtype=torch.float64)  # zero span, unsorted, a p == 0 event
    r = N.reshape_then_acc_count_only(flat_t, height=2, width=2)
    assert float(r[0, 0, 0]) == 1.0 and float(r[0, 0, 1]) == 2.0
    r2 = N.reshape_then_acc_count_pol(flat_t[:2], height=2, width=2)
    assert float(r2[0, 0, 0]) == 1.0 and float(r2[1, 0, 1]) == 1.0
    e = torch.zeros((0, 4), dtype=torch.float64)
    assert tuple(N.reshape_then_acc_count_only(e, height=3, width=5).shape) == (1, 3, 5) and not N.reshape_then_acc_count_pol(e, height=3, width=5).any()


@pytest.mark.parametrize("name", ["acc_count", "acc", "acc_count_pol", "acc_count_only", "acc_time", "acc_all", "acc_time_pol", "acc_exp",
                                  "flat", "flat_pol", "acc_intensity"])
def test_upstream_count_representations_match_the_reference(N, G, name):
    """imagenet.py:169-510 through one mixed-density launch each; counts and presence planes exact, normalised times
    (latest = max, earliest = min) to 1e-5 (2^30 time grid)"""
    import torch
    H, W = int(G["H"]), int(G["W"])
    rep = getattr(N, "reshape_then_" + name)(torch.tensor(G["events_s"].copy()), height=H, width=W)
    want = G["up_" + name]
    assert rep.dtype == torch.float32 and tuple(rep.shape) == want.shape
    assert_close(rep.numpy(), want, rtol=1e-5, atol=1e-7, what=name)
    if name in ("acc_count", "acc_count_pol", "acc_count_only", "acc_all", "flat", "flat_pol"):
        planes = {"acc_count": (0, 2), "acc_count_pol": (0, 1), "acc_count_only": (0,), "acc_all": (0, 1), "flat": (0,), "flat_pol": (0, 1)}[name]  # acc_intensity: a float32 quotient, held to 1e-5 above
        for c in planes:
            assert np.array_equal(rep.numpy()[c], want[c])


def test_upstream_acc_count_empty_sample(N):
    """imagenet.py:258-262: an empty sample is replaced by ten fake positive events at pixel (0, 0)"""
    import torch
    rep = N.reshape_then_acc_count(torch.zeros((0, 4), dtype=torch.float64), height=8, width=8)
    assert float(rep[0, 0, 0]) == 10.0 and float(rep[0].sum()) == 10.0 and float(rep[2].sum()) == 0.0
    assert abs(float(rep[1, 0, 0]) - 1.0) < 1e-6


def test_upstream_empty_samples(N):
    """imagenet.py:353-354 (zeros of the default size), :397-438 (empty index lists leave zeros), :483-486 (fake events)"""
    import torch
    e = torch.zeros((0, 4), dtype=torch.float64)
    assert tuple(N.reshape_then_acc_all(e, height=8, width=8).shape) == (6, N.IMAGE_H, N.IMAGE_W)
    assert tuple(N.reshape_then_flat(e, height=8, width=8).shape) == (1, 8, 8) and not N.reshape_then_flat_pol(e, height=8, width=8).any()
    rep = N.reshape_then_acc_time_pol(e, height=8, width=8)
    assert abs(float(rep[0, 0, 0]) - 1.0) < 1e-6 and float(rep.sum()) == float(rep[0, 0, 0])


def test_flat_ignores_timestamps(N, G):
    """reshape_then_flat* never read the time column: unsorted / constant stamps are fine"""
    import torch
    H, W = int(G["H"]), int(G["W"])
    ev = G["events_s"].copy()
    ev[:, 2] = 0.0
    assert np.array_equal(N.reshape_then_flat_pol(torch.tensor(ev), height=H, width=W).numpy(), G["up_flat_pol"])


def test_loader_table_matches_the_dataset_dispatch(N):
    """imagenet.py:1232-1272"""
    assert N.loader_for(None) is N.reshape_then_acc and N.loader_for("event_histogram") is N.reshape_then_acc_count_pol
    assert N.loader_for("timestamp_image") is N.reshape_then_acc_time_pol and N.loader_for("binary_event_image") is N.reshape_then_flat
    assert N.loader_for("reshape_then_optimized") is None and N.loader_for("no such loader") is None  # caller-side wrappers: not in this module
    assert N.loader_for("DiST") is N.reshape_then_acc_adj_sort and N.loader_for("sorted_time_surface") is N.reshape_then_acc_sort


# the rank-based loaders (imagenet.py:513-999) against the reference's own functions (oracle/gen_golden_nimagenet_sorted.py)
SORT_BASE = dict(neglect_polarity=False, global_time=True, strict=False, use_image=False, denoise_sort=False, denoise_image=False,
                 filter_flash=False, filter_noise=False, quantize_sort=None)
SORT_CASES = {
    "default": {}, "neglect": dict(neglect_polarity=True), "strict": dict(strict=True), "strict_neglect": dict(strict=True, neglect_polarity=True),
    "image": dict(use_image=True), "image_neglect_strict": dict(use_image=True, neglect_polarity=True, strict=True), "quant8": dict(quantize_sort=8),
    "quant_list": dict(quantize_sort=[4, 16], use_image=True), "quant_list_neglect": dict(quantize_sort=[4, 16], neglect_polarity=True, strict=True),
    "local_time": dict(global_time=False), "local_time_strict": dict(global_time=False, strict=True),
}


@pytest.fixture(scope="module")
def GS():
    return load(golden("nimg_sorted")[0][1])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("case", sorted(SORT_CASES))
def test_sorted_time_surface_matches_the_reference(N, GS, tag, case):
    """ranks, presence images and quantised copies are integers or ratios of small integers: bit exact"""
    import torch
    H, W = int(GS[tag + "_H"]), int(GS[tag + "_W"])
    rep = N.reshape_then_acc_sort(torch.tensor(GS[tag + "_events"].copy()), height=H, width=W, **{**SORT_BASE, **SORT_CASES[case]})
    want = GS[f"{tag}_sort_{case}"]
    assert rep.dtype == torch.float32 and tuple(rep.shape) == want.shape and not rep.is_cuda
    assert np.array_equal(rep.numpy(), want)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_dist_matches_the_reference(N, GS, tag):
    """DiST: the kernel's planes are exact (ranks mapped back to the float64 stamps), the image-domain steps are the reference's
    own float32 torch ops: bit exact"""
    import torch
    H, W = int(GS[tag + "_H"]), int(GS[tag + "_W"])
    rep = N.reshape_then_acc_adj_sort(torch.tensor(GS[tag + "_events"].copy()), height=H, width=W, **SORT_BASE)
    want = GS[tag + "_dist"]
    assert rep.dtype == torch.float32 and tuple(rep.shape) == want.shape
    assert np.array_equal(rep.numpy(), want), float(np.abs(rep.numpy() - want).max())


def test_sorted_loaders_edge_cases(N):
    """a polarity without events becomes one fake event at pixel (0, 0) (imagenet.py:641-646): strict gives a zero plane with the
    fake count visible in the image, non-strict raises like `hot.max()` on an empty tensor; the undefined denoise helper raises
    NameError as in the reference"""
    import torch
    ev = torch.tensor([[1.0, 2.0, 0.5, 1.0], [3.0, 0.0, 0.6, 1.0], [1.0, 2.0, 0.7, 1.0]], dtype=torch.float64)
    r = N.reshape_then_acc_sort(ev, height=4, width=4, **{**SORT_BASE, "strict": True, "use_image": True})
    assert tuple(r.shape) == (4, 4, 4) and float(r[0, 2, 1]) == 1.0 and float(r[1, 2, 1]) == 1.0 and float(r[1, 0, 3]) == 0.0
    assert float(r[2, 0, 0]) == 1.0 and float(r[2].sum()) == 1.0 and not r[3].any()
    with pytest.raises(RuntimeError):
        N.reshape_then_acc_sort(ev, height=4, width=4, **SORT_BASE)
    with pytest.raises(NameError):
        N.reshape_then_acc_sort(ev, height=4, width=4, **{**SORT_BASE, "denoise_sort": True, "strict": True})
