"""Writes tests/golden/img_*.npz: the reference's image pipeline on small seeded representations.

Runs only where /root/reference exists.  letterbox is the reference's own function (data_augment.py, imported by
path); resize_image / resize_image_process are dataset METHODS whose modules need h5py / torch_geometric (absent
offline), so their ten lines are restated in oracle/image_pipeline.py around the same cv2 calls and this script checks
that the oracle's letterbox equals the reference's before writing anything.
TEST INFRASTRUCTURE ONLY (fixture generator; nothing in the product package imports it)."""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import image_pipeline as oimg  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_data_augment", "/root/reference/ev-YOLOv6/yolov6/data/data_augment.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CASES = [  # name, H, W, C, img_size, mode, dtype
    ("gen1like_up_letterbox", 24, 30, 12, 64, "letterbox", np.float64),     # r > 1: INTER_LINEAR, pad top / bottom
    ("mpx_like_half_letterbox", 36, 64, 12, 32, "letterbox", np.float64),  # r = 0.5: INTER_AREA 2 x 2, pad
    ("mpx_like_squash", 36, 64, 12, 32, "squash", np.float64),             # precompute_reps: scale (2, 1.125), general INTER_AREA
    ("tall_letterbox", 50, 20, 12, 64, "letterbox", np.float32),            # H > W: pad left / right, upscale
    ("odd_down_letterbox", 45, 65, 12, 32, "letterbox", np.float32),       # non-integer shrink 2.03: general INTER_AREA, int() truncation
    ("c2_hist", 30, 38, 2, 48, "letterbox", np.float64),                    # C <= 4: cv2.resize on the whole image
    ("c5_voxel_squash", 33, 47, 5, 32, "squash", np.float32),
    ("same_size", 32, 32, 12, 32, "letterbox", np.float64),                 # r == 1: no resize at all
]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for i, (name, H, W, C, S, mode, dt) in enumerate(CASES):
        rng = np.random.default_rng(7000 + i)
        rep = (rng.random((H, W, C)) * (rng.random((H, W, C)) < 0.4)).astype(dt)  # sparse like a real representation
        rep[..., 0] -= 0.3 * (rng.random((H, W)) < 0.2)                             # some negative values (polarity channels)
        im = rep * 255
        if mode == "letterbox":
            res = oimg.resize_image(im, S)
            lb, ratio, pad = ref.letterbox(res, S, auto=False, scaleup=False)       # the reference's function
            mine = oimg.letterbox(res, S)
            assert np.array_equal(lb, mine), name
            img = lb
        else:
            img = oimg.resize_image_process(im, S)
        if img.ndim == 2:
            img = img[..., None]
        img = np.ascontiguousarray(img.transpose((2, 0, 1))[::-1])
        want = (img.astype(np.float32) / 255).astype(np.float32)
        assert np.array_equal(want, oimg.detector_input(rep, S, mode)), name
        np.savez_compressed(os.path.join(out_dir, f"img_{name}.npz"), rep=rep, img_size=np.int64(S), mode=np.array(mode), out=want)
        print(name, rep.shape, "->", want.shape)


def main_affine():
    """img_affine_*.npz: the reference's OWN random_affine (data_augment.py:110-150, get_transform_matrix drawing from a seeded
    `random`) on the letterboxed image, then the flips; the oracle's augmented_detector_input is checked against it first"""
    import random
    out_dir = os.path.join(ROOT, "tests", "golden")
    for i, (name, H, W, C, S, ud, lr) in enumerate([("gen1like", 24, 30, 12, 64, False, True), ("mpx_like", 36, 64, 12, 48, True, False),
                                                    ("c2", 30, 38, 2, 40, False, False), ("c5", 33, 47, 5, 56, True, True)]):
        rng = np.random.default_rng(7100 + i)
        rep = (rng.random((H, W, C)) * (rng.random((H, W, C)) < 0.4)).astype(np.float64)
        rep[..., 0] -= 0.3 * (rng.random((H, W)) < 0.2)
        lb = oimg.letterbox(oimg.resize_image(rep * 255, S, augment=True), S)
        random.seed(900 + i)
        M, _ = ref.get_transform_matrix(lb.shape[:2], (S, S), 10, 0.1, 10, 0.1)  # the hyper-parameters of configs/*_finetune: degrees, scale, shear, translate
        random.seed(900 + i)
        warped, _ = ref.random_affine(lb, (), degrees=10, translate=0.1, scale=0.1, shear=10, new_shape=(S, S))  # draws the same M
        img = warped
        if ud:
            img = np.flipud(img)
        if lr:
            img = np.fliplr(img)
        img = np.ascontiguousarray(img.transpose((2, 0, 1))[::-1])
        want = (img.astype(np.float32) / 255).astype(np.float32)
        assert np.array_equal(want, oimg.augmented_detector_input(rep, S, M, ud, lr)), name
        np.savez_compressed(os.path.join(out_dir, f"img_affine_{name}.npz"), rep=rep, img_size=np.int64(S), M=M, flip_ud=np.bool_(ud), flip_lr=np.bool_(lr), out=want)
        print("affine", name, rep.shape, "->", want.shape)


if __name__ == "__main__":
    main()
    main_affine()
