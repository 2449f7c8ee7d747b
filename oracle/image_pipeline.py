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
