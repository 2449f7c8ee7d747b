"""The oracle (oracle/*.py, numpy restatement) against fixtures produced by executing the reference
files themselves (oracle/gen_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import assert_close, golden, load
from oracle import gwd as ogwd
from oracle import representations as orep


def ids(cases):
    return [c[0] for c in cases]


ERGO = golden("ergo12_n*")


@pytest.mark.parametrize("name,path", ERGO, ids=ids(ERGO))
def test_ergo12(name, path):
    g = load(path)
    with np.errstate(all="ignore"):
        out = orep.ergo12(g["x"], g["y"], g["t"], g["p"], int(g["H"]), int(g["W"]))
    assert_close(out, g["out"], rtol=1e-12, atol=1e-15, what=name)


def test_ergo12_gen1_50k():
    g = load(golden("ergo12_gen1_50k")[0][1])
    out = orep.ergo12(g["x"], g["y"], g["t"], g["p"], 240, 304)
    assert_close(out.astype(np.float32), g["out"], rtol=1e-6, atol=1e-12, what="gen1")
    # integer-valued channels are bit exact: polarity sum (3), count sum (5), count means (2,4,7,11)
    for c in (2, 3, 4, 5, 7, 11):
        assert np.array_equal(out[:, :, c].astype(np.float32), g["out"][:, :, c])


def test_ergo12_f8_seconds():
    """n_imagenet structured f8 input with t in seconds: .astype(int64) truncation."""
    g = load(golden("ergo12_f8_seconds")[0][1])
    with np.errstate(all="ignore"):
        out = orep.ergo12(g["x"], g["y"], g["t_seconds"].astype(np.int64), g["p"], 30, 40)
    assert_close(out, g["out"], rtol=1e-12, atol=1e-15)


MDES = golden("mdes_*") + golden("mdmin_*")  # mdmin: the aggregation "min" (oracle/gen_golden_min.py)


@pytest.mark.parametrize("name,path", MDES, ids=ids(MDES))
def test_mixed_density_generic(name, path):
    g = load(path)
    with np.errstate(all="ignore"):
        out = orep.mixed_density_event_stack(g["x"], g["y"], g["t"], g["p"], int(g["H"]), int(g["W"]),
                                             g["win"].tolist(), g["func"].tolist(), g["agg"].tolist(), str(g["stacking"]))
    assert_close(out, g["out"], rtol=1e-12, atol=1e-15, what=name)


ES = golden("eventstack_*")


@pytest.mark.parametrize("name,path", ES, ids=ids(ES))
def test_event_stack(name, path):
    g = load(path)
    out = orep.event_stack(g["x"], g["y"], g["t"], (g["p"].astype(np.int32) + 1) // 2, int(g["H"]), int(g["W"]), 12)
    assert out.dtype == g["out"].dtype == np.float32
    assert np.array_equal(out, g["out"]), name  # integer valued: bit exact


TS = golden("timesurface_*")


@pytest.mark.parametrize("name,path", TS, ids=ids(TS))
def test_time_surface(name, path):
    g = load(path)
    p01 = ((g["p"].astype(np.int32) + 1) / 2).astype(np.int8)
    with np.errstate(all="ignore"):
        idx = orep.time_surface_indices(g["t"].astype(np.int32), 6)
    assert np.array_equal(idx, g["indices"])
    out = orep.time_surface(g["x"], g["y"], g["t"], p01, idx, int(g["H"]), int(g["W"]), tau=50000)
    assert_close(out, g["out"], rtol=1e-13, atol=0, what=name)


TG = golden("tore_gen1_*")


@pytest.mark.parametrize("name,path", TG, ids=ids(TG))
def test_tore_gen1(name, path):
    g = load(path)
    out = orep.tore_gen1(g["x"].astype(np.int32), g["y"].astype(np.int32), g["t"].astype(np.int32), g["p"].astype(np.int32), 6)
    assert out.dtype == np.float32 and out.shape == g["out"].shape
    assert_close(out, g["out"], rtol=3e-7, atol=0, what=name)


TF = golden("tore_fixed_*")


@pytest.mark.parametrize("name,path", TF, ids=ids(TF))
def test_tore_fixed(name, path):
    g = load(path)
    t = g["t"].astype(np.int32)
    out = orep.tore(g["x"].astype(np.int32) + 1, g["y"].astype(np.int32) + 1, t, g["p"].astype(np.int32), t[-1], int(g["k"]),
                    (int(g["H"]), int(g["W"])))
    assert_close(out, g["out"], rtol=3e-7, atol=0, what=name)


VT = golden("voxel_tonic_*")


@pytest.mark.parametrize("name,path", VT, ids=ids(VT))
def test_voxel_tonic(name, path):
    g = load(path)
    with np.errstate(all="ignore"):
        out = orep.voxel_tonic(g["x"], g["y"], g["t"].astype(np.int32), g["p"].astype(np.int32), int(g["H"]), int(g["W"]), 12)
    assert_close(out[:, None], g["out"], rtol=1e-12, atol=1e-15, what=name)


VS = golden("voxel_subpixel_*")


@pytest.mark.parametrize("name,path", VS, ids=[n for n, _ in VS])
def test_voxel_evlicious_subpixel(name, path):
    """divider > 1: float32 coordinates, 4-tap bilinear scatter (utils.py:93-103), fixtures from the reference's own code"""
    g = load(path)
    out = orep.voxel_evlicious(g["x"], g["y"], g["t"], g["p"], int(g["H"]), int(g["W"]), int(g["bins"]), bool(g["normalize"]), divider=int(g["divider"]))
    assert out.dtype == g["out"].dtype and out.shape == g["out"].shape
    assert np.array_equal(out, g["out"])  # the same float32 additions in the same order


VE = golden("voxel_evlicious_*")


@pytest.mark.parametrize("name,path", VE, ids=ids(VE))
def test_voxel_evlicious(name, path):
    g = load(path)
    out = orep.voxel_evlicious(g["x"], g["y"], g["t"], g["p"], int(g["H"]), int(g["W"]), int(g["bins"]), bool(g["normalize"]))
    assert out.dtype == np.float32
    if not bool(g["normalize"]):
        assert np.array_equal(out, g["out"])
    else:
        assert_close(out, g["out"], rtol=1e-6, atol=1e-7, what=name)


def test_voxel_gwd():
    g = load(golden("voxel_gwd_small")[0][1])
    out = orep.voxel_gwd(g["x"].astype(int), g["y"].astype(int), g["t01"], g["p"].astype(float), 40, 30, 5)
    assert_close(out, g["out"], rtol=1e-12, atol=1e-15)


def test_dispatch_branches():
    """get_item_transform (gen1_transforms.py:12-89) outputs, including the x255."""
    G = {n[len("dispatch_"):]: load(p) for n, p in golden("dispatch_*")}
    g = G["MixedDensityEventStack"]
    x, y, t, p, H, W = g["x"], g["y"], g["t"], g["p"], int(g["H"]), int(g["W"])
    assert_close(orep.ergo12(x, y, t, p, H, W) * 255, g["out"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(orep.event_stack(x, y, t, (p.astype(int) + 1) // 2, H, W) * 255, G["EventStack"]["out"])
    assert_close(orep.time_surface_gen1(x, y, t.astype(np.int32), p, H, W) * 255, G["ToTimesurface"]["out"], rtol=1e-13)
    assert_close(orep.tore_gen1(x.astype(np.int32), y.astype(np.int32), t.astype(np.int32), p.astype(np.int32)) * 255,
                 G["tore"]["out"], rtol=3e-7)
    vt = orep.voxel_tonic(x, y, t.astype(np.int32), p.astype(np.int32), H, W, 12).transpose(1, 2, 0) * 255
    assert_close(vt, G["ToVoxelGrid"]["out"], rtol=1e-12, atol=1e-12)
    img = orep.to_image(x, y, (p.astype(np.int32) + 1) // 2, H, W).transpose(1, 2, 0)
    img *= 255
    assert np.array_equal(img, G["ToImage"]["out"])


GA = golden("gwd_a_pair_*")


@pytest.mark.parametrize("name,path", GA, ids=ids(GA))
def test_gwd_a_pair(name, path):
    g = load(path)
    assert_close(ogwd.gwd_a_cost(g["Xs"], g["Xt"], float(g["h"])), g["out"], rtol=1e-9, what=name)


def test_otmi():
    g = load(golden("otmi_small")[0][1])
    ev = np.stack([g["x"], g["y"], g["t"], g["p"]], 1).astype(np.int32)
    c = ogwd.otmi(ev, g["rep"].astype(np.float64), int(g["H"]), int(g["W"]), int(g["rep_size"]))
    assert_close(c, g["out"], rtol=1e-6, what="otmi")


def test_window_bounds():
    for n in [0, 1, 2, 3, 7, 8, 9, 1000, 50000, 999999]:
        b = orep.sbn_window_bounds(n)
        assert b[0] == (0, n) and b[3][1] == 3 * (n // 3)
        assert b[4][0] == n // 2 and b[5][0] == n // 2 + n // 4 and b[6][0] == n // 2 + n // 4 + n // 8
        s = orep.event_stack_starts(n, 12)
        assert s[0] == 0 and all(s[i] <= s[i + 1] <= n for i in range(11))


# ---- post-representation image pipeline: oracle vs fixtures made with the reference's letterbox ----
IMG = [c for c in golden("img_*") if not c[0].startswith("img_affine")]
IMG_AFFINE = golden("img_affine_*")


@pytest.mark.parametrize("name,path", IMG_AFFINE, ids=[n for n, _ in IMG_AFFINE])
def test_oracle_augmented_image_pipeline_matches_the_reference_random_affine(name, path):
    """fixtures made with the reference's own random_affine (data_augment.py) and flips; the oracle runs cv2 around the same M"""
    from oracle import image_pipeline as oimg
    g = load(path)
    got = oimg.augmented_detector_input(g["rep"], int(g["img_size"]), g["M"], bool(g["flip_ud"]), bool(g["flip_lr"]))
    assert np.array_equal(got, g["out"])


def test_warp_affine_restatement_equals_cv2():
    """the description of cv::warpAffine the CUDA kernel implements (fixed-point coordinates, float weight table, double
    accumulation, border scalar indexed with channel & 3), restated in numpy, against cv2 itself: equal to the last bit"""
    import cv2
    from oracle import image_pipeline as oimg
    rng = np.random.default_rng(5)
    for k in range(5):
        H, W = 40 + 7 * k, 64 - 5 * k
        img = rng.random((H, W, 12)) * 255
        ang, sc = rng.uniform(-10, 10), rng.uniform(0.9, 1.1)
        M = cv2.getRotationMatrix2D((W / 2, H / 2), ang, sc)
        M[:, 2] += rng.uniform(-6, 6, 2)
        M[0, 1] += rng.uniform(-0.1, 0.1)
        want = cv2.warpAffine(img, M, dsize=(W + 3, H - 2), borderValue=(114, 114, 114))
        assert np.array_equal(oimg.warp_affine_restated(img, M, (W + 3, H - 2)), want)
        assert want[..., 3::4].min() == 0.0 or True  # every fourth channel is padded with 0 (checked below on a corner)
    far = np.array([[1.0, 0.0, 30.0], [0.0, 1.0, 0.0]])  # shifts the image right: the left columns are pure border
    out = cv2.warpAffine(np.ones((8, 8, 12)), far, dsize=(8, 8), borderValue=(114, 114, 114))
    assert out[0, 0].tolist() == [114.0, 114.0, 114.0, 0.0] * 3


@pytest.mark.parametrize("name,path", IMG, ids=[n for n, _ in IMG])
def test_image_pipeline_oracle_matches_reference_fixtures(name, path):
    from oracle import image_pipeline as oimg
    g = load(path)
    got = oimg.detector_input(g["rep"], int(g["img_size"]), str(g["mode"]))
    assert got.dtype == np.float32 and np.array_equal(got, g["out"])


# ---- EST learned quantisation layer: oracle vs fixtures made with the reference's classes ----
@pytest.mark.parametrize("name,path", golden("est_*"), ids=[n for n, _ in golden("est_*")])
def test_est_oracle_matches_reference_fixtures(name, path):
    import torch
    from oracle import est as oest
    g = load(path)
    ws = [g[f"w{i}"] for i in range(3)]
    bs = [g[f"b{i}"] for i in range(3)]
    with torch.no_grad():
        vox, lb = oest.est_forward(torch.tensor(g["events"]), ws, bs, tuple(int(v) for v in g["dim"]), int(g["image_size"]))
    assert np.allclose(lb.numpy(), g["out"], rtol=1e-6, atol=1e-6) and np.allclose(vox.numpy(), g["vox"], rtol=1e-6, atol=1e-6)


# ---- ev-licious stateful filters: oracle vs fixtures made with the reference's numba kernels ----
@pytest.mark.parametrize("name,path", golden("filter_*"), ids=[n for n, _ in golden("filter_*")])
def test_filter_oracle_matches_reference_fixtures(name, path):
    from oracle import filters as ofil
    g = load(path)
    H, W, n = int(g["H"]), int(g["W"]), len(g["x"])
    x, y, t, p = g["x"], g["y"], g["t"], g["p"]
    last = np.full((H, W), -np.inf)
    assert np.array_equal(ofil.refractory_period(np.ones(n, bool), x, y, t, float(g["refr_period"]), last), g["refr_mask"])
    assert np.array_equal(last, g["refr_state1"])
    act = np.zeros((H, W), np.int32)
    assert np.array_equal(ofil.contrast_threshold_control(act, np.zeros(n, bool), x, y, p, float(g["ctc_factor"])), g["ctc_mask"])
    assert np.array_equal(act, g["ctc_state1"])
    fx, fy = int(g["fx"]), int(g["fy"])
    m, cm = ofil.filter_events_resize(x, y, p, np.zeros(n, bool), np.zeros((H // fy, W // fx), np.float32), fx, fy)
    assert np.array_equal(m, g["rsz_mask"]) and np.array_equal(cm, g["rsz_state1"])
    for r in (1, 2, 3, 4):  # background activity (utils.py:169-178), whole stream in one go = the two pieces of the fixture
        ts = np.full((H, W), -np.inf)
        got = ofil.background_activity_filter(np.ones(n, bool), ts, x, y, t, float(g["ba_depth"]), r)
        assert np.array_equal(got, g[f"ba{r}_mask"]) and np.array_equal(ts, g[f"ba{r}_state1"])
