"""CPU execution of the EDLines device code (cube_slam_wu_b200/csrc/edlines_dev.cuh) against the oracle (oracle/oracle_edlines.cpp).

The EDLines stages are one-thread-per-item functions without shared memory or synchronisation; tests/emul/edlines_emul.cpp compiles the
same header with g++ and loops over the items, so these tests execute the code the CUDA kernels wrap -- bit for bit against the oracle --
without a GPU.  (The kernels' launch glue is covered by tests/test_edlines_gpu.py on a B200.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    if not os.path.exists(os.path.join(CUDA_INC, "vector_types.h")):
        pytest.skip("CUDA headers not found")
    so = str(tmp_path_factory.mktemp("emul") / "libedlines_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I", CUDA_INC,
                           os.path.join(HERE, "emul", "edlines_emul.cpp"), "-o", so])
    return C.CDLL(so)


def _run(emul, oracle, gray, filter=True, thres=15.0, cap=20000, descriptors=False):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    _, dx, dy = oracle.lbd_gradients(gray)
    out = np.zeros((cap, 4), np.float32); st = np.zeros(4, np.int64); nch = C.c_int()
    d32 = np.zeros((cap, 32), np.uint8); d72 = np.zeros((cap, 72), np.float32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n = emul.emul_edlines_detect(vp(dx), vp(dy), w, h, int(filter), C.c_float(thres), vp(out), cap, vp(st), C.byref(nch),
                                 vp(d32) if descriptors else None, vp(d72) if descriptors else None)
    if descriptors:
        return out[:n].copy(), d32[:n].copy(), d72[:n].copy()
    return out[:n].copy(), st, nch.value


def _same(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_device_code_matches_oracle_on_synthetic_frames(emul, oracle):
    from cube_slam_wu_b200 import synth
    total = 0
    for (w, h, seed, kw) in ((640, 480, 3, {}), (641, 479, 4, dict(texture=1.0, noise_sigma=5.0)), (1242, 375, 5, {}), (97, 61, 6, {})):
        gray = synth.make_lsd_frames(1, w, h, seed=seed, **kw)[0]
        for filt, thr in ((True, 15.0), (False, 0.0), (True, 50.0)):
            got, st, nch = _run(emul, oracle, gray, filt, thr)
            ref, _ = oracle.edlines_detect(gray, filter=filt, length_thres=thr)
            assert _same(got, ref), "%dx%d filter=%s: %d vs %d lines" % (w, h, filt, len(got), len(ref))
        chains = oracle.edlines_chains(gray)
        _, _, anchors = oracle.edlines_maps(gray)
        assert nch == len(chains) and st[2] == len(chains) and st[1] == sum(len(c) for c in chains) and st[0] == len(anchors)
        total += len(ref)
    assert total > 50


def test_device_code_on_the_reference_image_and_noise(emul, oracle):
    d = np.load(os.path.join(HERE, "golden", "lsd_407.npz"))
    got, st, nch = _run(emul, oracle, d["gray"])
    ref, _ = oracle.edlines_detect(d["gray"])
    assert _same(got, ref) and len(ref) > 100
    rng = np.random.default_rng(1)
    noise = rng.integers(0, 256, (120, 160), dtype=np.uint8)     # anchors everywhere, thousands of short walks
    got, st, nch = _run(emul, oracle, noise, False, 0.0)
    ref, _ = oracle.edlines_detect(noise, filter=False)
    assert _same(got, ref)
    flat = np.full((40, 60), 77, np.uint8)
    got, st, nch = _run(emul, oracle, flat)
    assert len(got) == 0 and nch == 0


def test_descriptor_device_code_matches_oracle(emul, oracle):
    """detect_descrip_lines with use_LSD = false: EDLines key lines (direction, chain-segment pixel count, projected end points) through the
    one-thread-per-item descriptor code (csrc/lbd_dev.cuh) against orc_lbd_describe_keylines -- 72 floats and 32 bytes, bit for bit."""
    from cube_slam_wu_b200 import synth
    frames = [synth.make_lsd_frames(1, 640, 480, seed=3)[0], synth.make_lsd_frames(1, 333, 251, seed=8, texture=1.0, noise_sigma=4.0)[0],
              np.load(os.path.join(HERE, "golden", "lsd_407.npz"))["gray"]]
    total = 0
    for gray in frames:
        lines, d32, d72 = _run(emul, oracle, gray, descriptors=True)
        ref_lines, extra = oracle.edlines_detect(gray)
        assert _same(lines, ref_lines)
        r72, r32 = oracle.lbd_describe_keylines(gray, ref_lines, extra[:, 0], extra[:, 1])
        assert np.array_equal(d32, r32)
        assert np.array_equal(np.isnan(d72), np.isnan(r72)) and np.array_equal(np.nan_to_num(d72).view(np.uint32), np.nan_to_num(r72).view(np.uint32))
        total += len(ref_lines)
    assert total > 200


def test_device_code_fuzz(emul, oracle):
    """random sizes and contents (synthetic scenes, textured + noisy scenes, contrast-stretched smooth noise, white noise): segments and
    descriptors from the device code stay bit-identical to the oracle's"""
    import cv2
    from cube_slam_wu_b200 import synth
    rng = np.random.default_rng(123)
    total = 0
    for it in range(16):
        w = int(rng.integers(40, 700)); h = int(rng.integers(40, 500))
        kind = it % 4
        if kind == 0:
            gray = synth.make_lsd_frames(1, w, h, seed=1000 + it)[0]
        elif kind == 1:
            gray = synth.make_lsd_frames(1, w, h, seed=1000 + it, texture=float(rng.uniform(0.5, 2)), noise_sigma=float(rng.uniform(2, 12)))[0]
        elif kind == 2:
            g = cv2.GaussianBlur(rng.normal(128, 60, (h, w)), (0, 0), float(rng.uniform(1, 4)))
            gray = np.clip((g - 128) * float(rng.uniform(2, 8)) + 128, 0, 255).astype(np.uint8)
        else:
            gray = rng.integers(0, 256, (h, w), dtype=np.uint8)
        filt = bool(it % 2)
        lines, d32, d72 = _run(emul, oracle, gray, filt, 15.0, descriptors=True)
        ref, extra = oracle.edlines_detect(gray, filter=filt, length_thres=15.0)
        assert _same(lines, ref), "case %d (%dx%d, kind %d)" % (it, w, h, kind)
        if len(ref):
            r72, r32 = oracle.lbd_describe_keylines(gray, ref, extra[:, 0], extra[:, 1])
            assert np.array_equal(d32, r32)
            assert np.array_equal(np.isnan(d72), np.isnan(r72)) and np.array_equal(np.nan_to_num(d72).view(np.uint32), np.nan_to_num(r72).view(np.uint32))
        total += len(ref)
    assert total > 200
