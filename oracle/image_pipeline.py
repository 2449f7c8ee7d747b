"""CPU ORACLE for the post-representation image pipeline.

TEST INFRASTRUCTURE ONLY (see oracle/representations.py for the rules): imported by tests/, bench_extra.py's CPU leg and
oracle/gen_golden_image.py, never by the product package.

Restates the reference's per-sample steps around OpenCV (cv2 is a third-party dependency of the reference and is
installed here, so cv2.resize / cv2.copyMakeBorder themselves are executed, not restated):
  resize_image          ev-YOLOv6/yolov6/data/gen1_2yolo.py:230-265
  resize_image_process  ev-YOLOv6/yolov6/data/gen4/precompute_reps.py:216-251
  letterbox             ev-YOLOv6/yolov6/data/data_augment.py:31-83 (auto=False, scaleup=False: the evaluation branch)
  CHW + reversal        gen1_2yolo.py:397
  / 255                 ev-YOLOv6/yolov6/core/engine.py:629-635
Pinned by tests/golden/img_*.npz, which oracle/gen_golden_image.py produces with the reference's OWN letterbox function
(imported from /root/reference) around the same cv2 calls.
"""
import numpy as np


def resize_image(im, img_size, augment=False):
    """gen1_2yolo.py:230-265"""
    import cv2
    h0, w0 = im.shape[:2]
    r = img_size / max(h0, w0)
    if r != 1:
        interp = cv2.INTER_AREA if r < 1 and not augment else cv2.INTER_LINEAR
        size = (int(w0 * r), int(h0 * r))
        if im.shape[2] > 4:
            im = cv2.merge([cv2.resize(c, size, interpolation=interp) for c in cv2.split(im)])
        else:
            im = cv2.resize(im, size, interpolation=interp)
    return im


def resize_image_process(im, img_size, augment=False):
    """precompute_reps.py:216-251 (squash to img_size x img_size)"""
    import cv2
    h0, w0 = im.shape[:2]
    r = img_size / max(h0, w0)
    if r != 1:
        interp = cv2.INTER_AREA if r < 1 and not augment else cv2.INTER_LINEAR
        size = (img_size, img_size)
        if im.shape[2] > 4:
            im = cv2.merge([cv2.resize(c, size, interpolation=interp) for c in cv2.split(im)])
        else:
            im = cv2.resize(im, size, interpolation=interp)
    return im


def letterbox(im, new_shape, color=114.0):
    """data_augment.py:31-83 with auto=False, scaleup=False; after resize_image the image already fits, so r == 1"""
    import cv2
    shape = im.shape[:2]
    r = min(new_shape / shape[0], new_shape / shape[1])
    r = min(r, 1.0)
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = (new_shape - new_unpad[0]) / 2, (new_shape - new_unpad[1]) / 2
    if shape[::-1] != new_unpad:
        im = cv2.merge([cv2.resize(c, new_unpad, interpolation=cv2.INTER_LINEAR) for c in cv2.split(im)])
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return cv2.merge([cv2.copyMakeBorder(c, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color) for c in cv2.split(im)])


def detector_input(rep, img_size, mode="letterbox"):
    """rep (H, W, C) as the representation returns it (before x255) -> float32 (C, img_size, img_size) as the model sees it"""
    im = np.asarray(rep) * 255  # get_item_transform's final x255 (gen1_transforms.py)
    if mode == "letterbox":
        im = letterbox(resize_image(im, img_size), img_size)
    else:
        im = resize_image_process(im, img_size)
    if im.ndim == 2:
        im = im[..., None]
    im = np.ascontiguousarray(im.transpose((2, 0, 1))[::-1])
    return (im.astype(np.float32) / 255).astype(np.float32)  # .float() / 255


def warp_affine_restated(img, M23, dsize, border=(114.0, 114.0, 114.0, 0.0)):
    """cv::warpAffine(INTER_LINEAR, BORDER_CONSTANT) on a float image, restated from imgwarp.cpp (WarpAffineInvoker +
    remapBilinear): inverse map in double, destination coordinates in fixed point (10 fractional bits, rounded to 1/32 pixel),
    float 32 x 32 weight table, double accumulation, border scalar indexed with channel & 3.  Used to CHECK that description
    against cv2 itself (tests/test_oracle_golden.py: equal to the last bit on float64 images); the oracle proper calls cv2."""
    H, W, C = img.shape
    ow, oh = dsize
    M = np.array(M23, np.float64).reshape(6).copy()
    D = M[0] * M[4] - M[1] * M[3]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[4] * D, M[0] * D
    M[0] = A11
    M[1] *= -D
    M[3] *= -D
    M[4] = A22
    b1 = -M[0] * M[2] - M[1] * M[5]
    b2 = -M[3] * M[2] - M[4] * M[5]
    M[2], M[5] = b1, b2
    x, y = np.arange(ow, dtype=np.float64), np.arange(oh, dtype=np.float64)
    rnd = lambda v: np.rint(v).astype(np.int64)  # noqa: E731  saturate_cast<int>(double): round half to even
    adelta, bdelta = rnd(M[0] * x * 1024), rnd(M[3] * x * 1024)
    X0, Y0 = rnd((M[1] * y + M[2]) * 1024) + 16, rnd((M[4] * y + M[5]) * 1024) + 16
    X, Y = (X0[:, None] + adelta[None, :]) >> 5, (Y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
    f32 = np.float32
    fx, fy = (X & 31).astype(f32) / f32(32), (Y & 31).astype(f32) / f32(32)
    w = [(f32(1) - fy) * (f32(1) - fx), (f32(1) - fy) * fx, fy * (f32(1) - fx), fy * fx]
    bv = np.array([border[k & 3] for k in range(C)], np.float64)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = img[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok[..., None], v, bv[None, None, :])
    return (tap(sy, sx) * w[0][..., None].astype(np.float64) + tap(sy, sx + 1) * w[1][..., None] + tap(sy + 1, sx) * w[2][..., None]
            + tap(sy + 1, sx + 1) * w[3][..., None])


def augmented_detector_input(rep, img_size, M, flip_ud=False, flip_lr=False):
    """The training branch of Gen1H5.__getitem__ (gen1_2yolo.py:321-397) for given random draws: x255, resize_image with
    INTER_LINEAR (augment=True), letterbox, random_affine's cv2.warpAffine(img, M[:2], dsize=(img_size, img_size),
    borderValue=(114, 114, 114)) (data_augment.py:110-123), general_augment's flips, CHW + reversal, / 255"""
    import cv2
    im = np.asarray(rep) * 255
    im = letterbox(resize_image(im, img_size, augment=True), img_size)
    M = np.asarray(M, np.float64)
    if (M != np.eye(3)).any():
        im = cv2.warpAffine(im, M[:2], dsize=(img_size, img_size), borderValue=(114, 114, 114))
    if flip_ud:
        im = np.flipud(im)
    if flip_lr:
        im = np.fliplr(im)
    im = np.ascontiguousarray(im.transpose((2, 0, 1))[::-1])
    return (im.astype(np.float32) / 255).astype(np.float32)
