#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/.

Run in the build container only (it reads /root/reference and imports cv2 4.13.0):

    python tests/golden/make_golden.py

What it pins, and against what:
  * brief_offsets.npy        the fixed 256x4 BRIEF offset table: the reference draws
                             it from std::random_device (src/BriefDescriptor.cc:4-20), so tests
                             and benchmarks inject this one.
  * kitti_frame.png          the real 376x1241 8-bit KITTI-shaped frame decoded from the
                             reference's tests/epilinesOpencv.png (input pixels only).
  * bresenham_50x50.npy      decoded tests/testBresenham.png (the reference's golden ring image,
                             tests/FastDetectorTest.cc:6-31).
  * blur_golden.npz          cv2.GaussianBlur(img, (9,9), 2.5, 2.5) outputs: full arrays for small
                             images, sha256 for seeded full-size ones (src/BriefDescriptor.cc:90).
  * eigen_golden.npz         cv2.eigen on 2x2 float32 structure tensors (src/FastDetector.cc:265)
                             and the Harris response of src/FastDetector.cc:270 built on them.
  * ref_golden.npz           outputs of the reference's OWN FastDetector/Brief sources compiled
                             unmodified over the stub OpenCV tree (oracle/_ref, if built).
"""
import hashlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synth_frame(kind, seed, H=376, W=1241):
    """Synthetic inputs of SURVEY 8d (also used by tests/ and bench.py via ya_vo_b200.synth)."""
    from ya_vo_b200.synth import synth_frame as sf
    return sf(kind, seed, H, W)


def main():
    from oracle import pyoracle as po
    po.build()

    # ---- fixed BRIEF table -------------------------------------------------------------
    offsets = np.random.default_rng(7).integers(-8, 9, (256, 4)).astype(np.int32)
    np.save(os.path.join(HERE, "brief_offsets.npy"), offsets)

    # ---- input fixtures decoded from the reference's test images ------------------------
    kitti = cv2.imread(os.path.join(REF, "tests/epilinesOpencv.png"), 0)
    assert kitti.shape == (376, 1241) and kitti.dtype == np.uint8
    cv2.imwrite(os.path.join(HERE, "kitti_frame.png"), kitti, [cv2.IMWRITE_PNG_COMPRESSION, 9])
    assert np.array_equal(cv2.imread(os.path.join(HERE, "kitti_frame.png"), 0), kitti)
    bres = cv2.imread(os.path.join(REF, "tests/testBresenham.png"), 0)
    np.save(os.path.join(HERE, "bresenham_50x50.npy"), bres)

    # ---- GaussianBlur ---------------------------------------------------------------------
    blur = {}
    rng = np.random.default_rng(11)
    small = {
        "small_50x37": rng.integers(0, 256, (37, 50), dtype=np.uint8),
        "small_9x9": rng.integers(0, 256, (9, 9), dtype=np.uint8),
        "small_5x64": rng.integers(0, 256, (5, 64), dtype=np.uint8),
        "small_64x5": rng.integers(0, 256, (64, 5), dtype=np.uint8),
        "small_sat": np.full((20, 33), 255, np.uint8),
        "kitti_crop": kitti[100:164, 600:700].copy(),
    }
    for name, img in small.items():
        out = cv2.GaussianBlur(img, (9, 9), 2.5, None, 2.5)
        assert np.array_equal(out, po.gaussian_blur(img)), name
        blur[name + "_in"] = img
        blur[name + "_out"] = out
    big = {"U_seed0": synth_frame("U", 0), "G30_seed1": synth_frame("G30", 1), "B4_seed2": synth_frame("B4", 2),
           "kitti": kitti, "U4K_seed3": synth_frame("U", 3, 2160, 3840)}
    for name, img in big.items():
        for opt in (True, False):
            cv2.setUseOptimized(opt)
            out = cv2.GaussianBlur(img, (9, 9), 2.5, None, 2.5)
            assert np.array_equal(out, po.gaussian_blur(img)), (name, opt)
        cv2.setUseOptimized(True)
        blur[name + "_sha_in"] = np.array(sha(img))
        blur[name + "_sha_out"] = np.array(sha(out))
    np.savez_compressed(os.path.join(HERE, "blur_golden.npz"), **blur)
    print("blur: oracle == cv2.GaussianBlur on", len(small) + len(big), "images")

    # ---- cv::eigen 2x2 float32 + Harris response ------------------------------------------
    rng = np.random.default_rng(5)
    tensors = []
    # realistic: 3x3 box sums of Sobel products on real / synthetic pixels
    for img in (kitti, synth_frame("U", 0), synth_frame("G30", 1)):
        ix, iy = po.sobel(img)
        ix = ix.astype(np.int64)
        iy = iy.astype(np.int64)
        ys = rng.integers(4, img.shape[0] - 4, 1500)
        xs = rng.integers(4, img.shape[1] - 4, 1500)
        for r, c in zip(ys, xs):
            wx = ix[r - 1:r + 2, c - 1:c + 2]
            wy = iy[r - 1:r + 2, c - 1:c + 2]
            tensors.append(((wx * wx).sum(), (wx * wy).sum(), (wy * wy).sum()))
    # adversarial: zero / tiny off-diagonal, equal diagonal, extremes
    for a, b, c in [(0, 0, 0), (5, 0, 3), (3, 0, 5), (7, 7, 7), (1, 1, 1), (9363600, 9363600, 9363600),
                    (9363600, -9363600, 9363600), (9363600, 0, 0), (0, 1, 0), (100, -1, 100), (2, 1, 1),
                    (1, 1, 2), (16777215, 1, 0), (123456, -654321, 7654321)]:
        tensors.append((a, b, c))
    t = np.array(tensors, np.float32)
    l = np.zeros((t.shape[0], 2), np.float32)
    score = np.zeros(t.shape[0], np.float32)
    for i, (a, b, c) in enumerate(t):
        M = np.array([[a, b], [b, c]], np.float32)
        ok, ev = cv2.eigen(M)[:2]
        l[i] = ev.reshape(-1)
        l1, l2 = np.float32(ev[0, 0]), np.float32(ev[1, 0])
        # src/FastDetector.cc:270: float product, double pow/sub, narrowed to float
        score[i] = np.float32(np.float64(np.float32(l1 * l2)) - 0.04 * np.float64(np.float32(l2 + l1)) ** 2)
        o1, o2 = po.eigen2x2(a, b, c)
        assert o1.tobytes() == l[i, 0].tobytes() and o2.tobytes() == l[i, 1].tobytes(), (a, b, c, o1, o2, l[i])
        assert po.score_from_tensor(a, b, c).tobytes() == score[i].tobytes(), (a, b, c)
    np.savez_compressed(os.path.join(HERE, "eigen_golden.npz"), tensors=t, eig=l, score=score)
    print("eigen: oracle == cv2.eigen on", t.shape[0], "tensors")

    # ---- reference sources compiled over the shim (optional) -------------------------------
    try:
        from oracle import pyref
        pyref.write_golden(os.path.join(HERE, "ref_golden.npz"))
    except Exception as e:  # noqa
        print("ref_golden.npz not regenerated:", e)


if __name__ == "__main__":
    main()
