"""The fused image pipeline (rep -> x255 -> cv2.resize -> letterbox -> CHW reversed -> /255) on the GPU against the
fixtures produced with the reference's own letterbox + cv2 (tests/golden/img_*.npz) and against the cv2 oracle at the
BASELINE sizes.  Bar: 1e-5 relative (north_star) with an absolute floor of 2e-6: the reference resizes a float64 image
with float32 weights, the kernel a float32 image, and outputs near zero are sums of cancelling positive and negative taps
(values are O(1) after the / 255)."""
import numpy as np
import pytest

from conftest import assert_close, golden, load

pytestmark = pytest.mark.gpu
ATOL = 2e-6


@pytest.fixture(scope="module")
def E(cuda_device):
    import event_representation_study_b200.batched as eb
    return eb


IMG = [c for c in golden("img_*") if not c[0].startswith("img_affine")]
IMG_AFFINE = golden("img_affine_*")


@pytest.mark.parametrize("name,path", IMG_AFFINE, ids=[n for n, _ in IMG_AFFINE])
def test_augmented_pipeline_matches_the_reference_random_affine(E, name, path):
    """training branch (gen1_2yolo.py:321-397): resize (INTER_LINEAR) + letterbox on the GPU, then evrep_warp_affine_batched with
    the M the reference's get_transform_matrix drew and the flips, against the reference's own random_affine output"""
    import torch
    g = load(path)
    S = int(g["img_size"])
    rep = torch.as_tensor(g["rep"].astype(np.float32)).cuda()[None]
    lb = E.detector_input(rep, S, mode="letterbox", interp="linear", scale_out=1.0, reverse_channels=False)
    got = E.augment_affine(lb, g["M"][None], [bool(g["flip_ud"])], [bool(g["flip_lr"])])[0].cpu().numpy()
    assert got.shape == g["out"].shape
    assert_close(got, g["out"], rtol=1e-5, atol=ATOL, what=name)


def test_augmented_pipeline_at_detector_size_vs_cv2_oracle(E):
    """Gen1 representation -> 640 x 640 with four different draws (one the identity: the reference skips the warp then), batched"""
    import math
    import random
    import torch
    from oracle import image_pipeline as oimg
    H, W, S, B = 240, 304, 640, 4
    rng = np.random.default_rng(11)
    reps = (rng.random((B, H, W, 12)) * (rng.random((B, H, W, 12)) < 0.3)).astype(np.float32)
    random.seed(3)
    Ms = []
    for b in range(B):
        a, s_ = random.uniform(-10, 10), random.uniform(0.9, 1.1)
        Cm, R, Sh, T = np.eye(3), np.eye(3), np.eye(3), np.eye(3)
        Cm[0, 2], Cm[1, 2] = -S / 2, -S / 2
        ca, sa = math.cos(math.radians(a)) * s_, math.sin(math.radians(a)) * s_
        R[:2] = [[ca, sa, 0.0], [-sa, ca, 0.0]]  # cv2.getRotationMatrix2D(angle=a, center=(0, 0), scale=s)
        Sh[0, 1], Sh[1, 0] = math.tan(math.radians(random.uniform(-10, 10))), math.tan(math.radians(random.uniform(-10, 10)))
        T[0, 2], T[1, 2] = random.uniform(0.4, 0.6) * S, random.uniform(0.4, 0.6) * S
        Ms.append(T @ Sh @ R @ Cm if b else np.eye(3))
    ud, lr = [False, True, False, True], [False, False, True, True]
    lb = E.detector_input(torch.as_tensor(reps).cuda(), S, interp="linear", scale_out=1.0, reverse_channels=False)
    got = E.augment_affine(lb, np.stack(Ms), ud, lr).cpu().numpy()
    for b in range(B):
        assert_close(got[b], oimg.augmented_detector_input(reps[b], S, Ms[b], ud[b], lr[b]), rtol=1e-5, atol=ATOL, what=f"window {b}")


@pytest.mark.parametrize("name,path", IMG, ids=[n for n, _ in IMG])
def test_image_pipeline_matches_reference_fixtures(E, name, path):
    import torch
    g = load(path)
    rep = torch.as_tensor(g["rep"].astype(np.float32)).cuda()[None]
    got = E.detector_input(rep, int(g["img_size"]), mode=str(g["mode"]))[0].cpu().numpy()
    assert_close(got, g["out"], rtol=1e-5, atol=ATOL, what=name)


@pytest.mark.parametrize("H,W,S,mode", [(240, 304, 640, "letterbox"), (720, 1280, 640, "letterbox"), (720, 1280, 640, "squash"),
                                       (240, 304, 224, "letterbox")])
def test_image_pipeline_baseline_sizes_vs_cv2_oracle(E, H, W, S, mode):
    import torch
    from oracle import image_pipeline as oimg
    rng = np.random.default_rng(H + S)
    reps = (rng.random((3, H, W, 12)) * (rng.random((3, H, W, 12)) < 0.3)).astype(np.float32)
    reps[..., 3] -= 0.5 * (rng.random((3, H, W)) < 0.2)
    got = E.detector_input(torch.as_tensor(reps).cuda(), S, mode=mode).cpu().numpy()
    assert got.shape == (3, 12, S, S)
    for b in range(3):
        assert_close(got[b], oimg.detector_input(reps[b], S, mode), rtol=1e-5, atol=ATOL, what=f"window {b}")


def test_image_pipeline_generic_channels_and_options(E):
    import torch
    from oracle import image_pipeline as oimg
    rng = np.random.default_rng(3)
    rep = rng.random((2, 60, 80, 5)).astype(np.float32)
    t = torch.as_tensor(rep).cuda()
    got = E.detector_input(t, 96, mode="letterbox").cpu().numpy()
    for b in range(2):
        assert_close(got[b], oimg.detector_input(rep[b], 96, "letterbox"), rtol=1e-5, atol=ATOL)
    # no reversal, no scaling, another pad value: the raw resize + letterbox
    raw = E.detector_input(t, 96, scale_in=1.0, scale_out=1.0, pad_value=0.5, reverse_channels=False).cpu().numpy()
    want = oimg.letterbox(oimg.resize_image(rep[0].astype(np.float64), 96), 96, color=0.5).transpose(2, 0, 1)
    assert_close(raw[0], want, rtol=1e-5, atol=ATOL)


def test_image_pipeline_feeds_from_the_representation(E):
    """end to end on the device: events -> ERGO-12 -> detector input, against oracle representation + cv2 pipeline"""
    from event_representation_study_b200.synth import poisson_window
    from oracle import image_pipeline as oimg
    from oracle import representations as orep
    H, W = 240, 304
    w = poisson_window(77, 30000, H, W)
    ev = E.pack_events([w], "cuda")
    img = E.detector_input(E.ergo12(ev, H, W), 640).cpu().numpy()[0]
    want = oimg.detector_input(np.nan_to_num(orep.ergo12(w["x"], w["y"], w["t"], w["p"], H, W)), 640)
    assert_close(np.nan_to_num(img), want, rtol=1e-5, atol=5e-6)


def test_image_pipeline_rejects_what_it_does_not_mirror(E):
    import torch
    with pytest.raises(ValueError):
        E.detector_input(torch.zeros((1, 8, 8, 12)), 16)


@pytest.mark.parametrize("H,W,S,mode", [(64, 2000, 64, "letterbox"), (720, 1280, 96, "squash"), (720, 1280, 100, "letterbox"), (333, 517, 31, "squash")])
def test_image_pipeline_large_inter_area_scales(E, H, W, S, mode):
    """INTER_AREA at any shrink factor (round 1 stopped at 4): taps are generated on the fly, 31 x per axis here"""
    import torch
    from oracle import image_pipeline as oimg
    rng = np.random.default_rng(H + W + S)
    rep = (rng.random((2, H, W, 12)) * (rng.random((2, H, W, 12)) < 0.5)).astype(np.float32)
    got = E.detector_input(torch.as_tensor(rep).cuda(), S, mode=mode).cpu().numpy()
    for b in range(2):
        assert_close(got[b], oimg.detector_input(rep[b], S, mode), rtol=1e-5, atol=ATOL, what=f"window {b}")


def test_image_pipeline_empty_batch_and_single_pixel(E):
    import torch
    out = E.detector_input(torch.zeros((0, 8, 8, 12), device="cuda"), 16)
    assert tuple(out.shape) == (0, 12, 16, 16)
    one = E.detector_input(torch.full((1, 1, 1, 12), 0.5, device="cuda"), 4, scale_in=1.0, scale_out=1.0).cpu().numpy()
    assert np.allclose(one, 0.5)  # a 1 x 1 image enlarged to 4 x 4: every tap is the single pixel, no border
