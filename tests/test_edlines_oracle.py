"""CPU tests of the EDLines oracle (oracle/oracle_edlines.cpp): the use_LSD = false branch of line_lbd_detect, which object_slam selects
(main_obj.cpp:503-505).  There is no GPU path for it yet; the oracle fixes the semantics a port will be held to.

PARITY UNPINNED for the segment lists (the reference ships no EDLines output; cv2 4.13 has no EDLines).  Checked here: the per-pixel maps
and the anchors against an independent numpy restatement on top of the cv2-pinned blur / Sobel, structural invariants of the edge chains
and of the fitted segments, and the relation to the LSD branch on the reference's own image."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _frame(seed=3, w=320, h=240):
    from cube_slam_wu_b200 import synth
    return synth.make_lsd_frames(1, w, h, seed=seed)[0]


def test_maps_and_anchors_match_numpy(oracle):
    gray = _frame()
    g, d, anchors = oracle.edlines_maps(gray)
    _, dx, dy = oracle.lbd_gradients(gray)          # blur 5x5 + Sobel, pinned against cv2 in test_lbd_oracle.py
    ax, ay = np.abs(dx.astype(np.int32)), np.abs(dy.astype(np.int32))
    s = ax + ay
    ref_g = np.rint(np.where(s > 81, s, 0) * 0.25).astype(np.int16)   # threshold TOZERO at 80 + 1, MatExpr / 4 = round half to even
    assert np.array_equal(g, ref_g)
    assert np.array_equal(d, np.where(ax < ay, 255, 0).astype(np.uint8))
    assert 0 < g.max() < 256                                          # after the 5x5 blur the routing's uchar casts never wrap in practice
    # anchors: column by column, every second pixel, local maxima by 8 across the edge direction (binary_descriptor.cpp:1643-1670)
    gi = g.astype(np.int32)
    h, w = g.shape
    ref = []
    for x in range(1, w - 1, 2):
        for y in range(1, h - 1, 2):
            if d[y, x] == 255:
                ok = gi[y, x] >= gi[y - 1, x] + 8 and gi[y, x] >= gi[y + 1, x] + 8
            else:
                ok = gi[y, x] >= gi[y, x - 1] + 8 and gi[y, x] >= gi[y, x + 1] + 8
            if ok:
                ref.append((x, y))
    assert np.array_equal(anchors, np.array(ref, np.int64).reshape(-1, 2))
    assert len(ref) > 200


def test_edge_chains_invariants(oracle):
    gray = _frame(seed=5)
    g, d, anchors = oracle.edlines_maps(gray)
    chains = oracle.edlines_chains(gray)
    assert len(chains) > 10
    seen = set()
    for c in chains:
        assert len(c) >= 15                                   # first + second part >= minLineLen + 1, the anchor counted once
        step = np.abs(np.diff(c, axis=0))
        assert step.max() <= 1 and (step.sum(axis=1) > 0).all()   # 8-connected walk, no pixel repeated in place
        assert (g[c[:, 1], c[:, 0]] > 0).all()                # chains stay inside the thresholded gradient
        for p in map(tuple, c):
            assert p not in seen                              # the edge map stops every walk at a drawn pixel
            seen.add(p)
    anchor_set = set(map(tuple, anchors))
    assert sum(1 for c in chains if anchor_set & set(map(tuple, c))) == len(chains)   # every chain was started from an anchor


def test_segments_are_consistent(oracle):
    gray = _frame(seed=7, w=640, h=480)
    lines, extra = oracle.edlines_detect(gray)
    raw, raw_extra = oracle.edlines_detect(gray, filter=False)
    assert 20 < len(lines) <= len(raw)
    assert np.array_equal(lines, raw[raw_extra[:, 2] > 15.0])                  # filter_lines: lineLength > line_length_thres, order kept
    v = lines[:, 2:] - lines[:, :2]
    assert np.allclose(np.hypot(v[:, 0], v[:, 1]), extra[:, 2], rtol=1e-5)      # lineLength
    ang = np.arctan2(v[:, 1], v[:, 0])
    diff = np.abs(np.angle(np.exp(1j * (ang - extra[:, 0]))))
    # end points ordered along lineDirection_ (dark side on the left).  LineValidation_ builds the direction from |a|, |b| of the line
    # equation and the quadrant of the mean gradient (:2826-2841), so for a nearly axis-parallel line it can be the mirror image of the
    # fitted slope: exact for most segments, off by twice a small slope for a few
    assert np.median(diff) < 1e-5 and diff.max() < 0.2
    assert (extra[:, 1] >= 15).all() and (extra[:, 1] + 1 >= extra[:, 2] / np.sqrt(2) - 1).all()   # numOfPixels vs length of an 8-connected chain
    assert lines.min() > -2 and lines[:, 0::2].max() < 642 and lines[:, 1::2].max() < 482
    # a longer threshold keeps a subset
    long_lines, _ = oracle.edlines_detect(gray, length_thres=50.0)
    assert 0 < len(long_lines) < len(lines)
    # deterministic; the specified atan2 and libm's give the same segments here (they differ by <= 1 ulp and only feed comparisons / one float)
    again, _ = oracle.edlines_detect(gray)
    assert np.array_equal(again, lines)
    lit, lit_extra = oracle.edlines_detect(gray, libm_trig=True)
    assert np.array_equal(lit, lines) and np.abs(lit_extra[:, 0] - extra[:, 0]).max() <= 2.4e-7


def test_reference_image_edlines_vs_lsd(oracle):
    """line_lbd/data/407.jpg (fixture lsd_407.npz): EDLines with the reference's gradient threshold 80 finds the strong edges only -- fewer
    segments than its LSD branch, most of them close to an LSD segment (mid-point within 3 px of an LSD segment's support line and
    within its extent)."""
    d = np.load(os.path.join(GOLD, "lsd_407.npz"))
    ed, extra = oracle.edlines_detect(d["gray"])
    lsd = d["ref_lines"]
    assert 100 < len(ed) < len(lsd)
    mid = 0.5 * (ed[:, :2] + ed[:, 2:])
    p, q = lsd[:, :2], lsd[:, 2:]
    u = q - p
    L = np.linalg.norm(u, axis=1)
    u = u / L[:, None]
    rel = mid[:, None, :] - p[None, :, :]
    t = (rel * u[None]).sum(-1)
    dist = np.abs(rel[..., 0] * u[None, :, 1] - rel[..., 1] * u[None, :, 0])
    near = ((dist < 3.0) & (t > -5) & (t < L[None] + 5)).any(axis=1)
    assert near.mean() > 0.6, near.mean()


def test_tiny_and_flat_frames(oracle):
    flat = np.full((40, 60), 128, np.uint8)
    lines, _ = oracle.edlines_detect(flat)
    assert len(lines) == 0 and len(oracle.edlines_chains(flat)) == 0
    rng = np.random.default_rng(0)
    noise = rng.integers(0, 256, (33, 47), dtype=np.uint8)
    oracle.edlines_detect(noise, filter=False)   # must not crash on a frame full of anchors


def test_edlines_keylines_feed_the_descriptor_oracle(oracle):
    """detect_descrip_lines with use_LSD = false: the EDLines key lines (direction = lineDirection_, numOfPixels = chain-segment pixels) go
    through computeLBD unchanged.  Checked here on the CPU only (the GPU descriptor kernel takes LSD-style key lines so far): unit-norm
    descriptors, and -- for a line whose EDLines fields coincide with the LSD recipe's -- the same descriptor as the LSD-style entry point."""
    d = np.load(os.path.join(GOLD, "lsd_407.npz"))
    lines, extra = oracle.edlines_detect(d["gray"])
    d72, d32 = oracle.lbd_describe_keylines(d["gray"], lines, extra[:, 0], extra[:, 1])
    ok = ~np.isnan(d72).any(axis=1)
    assert ok.mean() > 0.95
    assert np.allclose(np.linalg.norm(d72[ok].astype(np.float64), axis=1), 1.0, atol=1e-5)
    assert len(np.unique(d32, axis=0)) > 0.9 * len(d32)
    # the LSD-style entry point derives angle = atan2(dy, dx) and numOfPixels = max(|dx|, |dy|) + 1 from the end points: where those agree with
    # the EDLines fields the two entry points must give the same bits
    r72, r32, kl = oracle.lbd_describe(d["gray"], lines)
    same_fields = (kl[:, 0] == extra[:, 0]) & (kl[:, 1] == extra[:, 1])
    if same_fields.any():
        assert np.array_equal(d32[same_fields], r32[same_fields])
    # a good part of the lines differ in numOfPixels (chain pixels vs LineIterator count; 29 % on this image), which is why the EDLines chain
    # needs the detector's fields
    assert (kl[:, 1] != extra[:, 1]).mean() > 0.1
