"""Pins oracle/gt_heatmap_oracle.py (ground-truth heatmap synthesis feeding the lifting
net under --use_gt_heatmap):
  * the Gaussian filter restatement against scipy.ndimage.gaussian_filter (installed; the reference's own dependency)
  * line_aa against the known-answer example of skimage's docstring (scikit-image itself is absent: published algorithm
    restated, see the oracle header)
  * coord2d_to_heatmap / get_limb_data / the loader's scaling + cos/sin modulation + the wrapper's channel order against
    the unmodified reference functions, live (skipped where the reference tree is absent), with the oracle's line_aa
    injected for the missing skimage import
  * the committed golden written from that live run (tests/golden/make_gt_heatmap_golden.py)"""
import os
import sys
import types

import numpy as np
import pytest

import gt_heatmap_oracle as gto
import ref_shim

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_gt_heatmaps.npz")


def test_gaussian_blur_matches_scipy():
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(0)
    img = rng.random((72, 72)).astype(np.float32)
    for mode in ("reflect", "constant"):
        ref = gaussian_filter(img, sigma=1.0, mode=mode)
        assert np.abs(gto.gaussian_blur(img, 1.0, mode) - ref).max() < 2e-7


def test_line_aa_known_answer_from_the_skimage_docstring():
    img = np.zeros((10, 10), dtype=np.uint8)
    rr, cc, val = gto.line_aa(1, 1, 8, 8)
    img[rr, cc] = val * 255
    want = np.zeros((10, 10), dtype=np.uint8)
    for i in range(1, 9):
        want[i, i] = 255
        if i < 8:
            want[i, i + 1] = want[i + 1, i] = 74
    assert np.array_equal(img, want)
    # structural properties of the published algorithm: endpoints are hit, axis-aligned lines are solid
    for a, b in (((3, 4), (20, 9)), ((20, 9), (3, 4)), ((5, 30), (5, 2)), ((0, 0), (0, 0)), ((7, 1), (40, 1))):
        rr, cc, val = gto.line_aa(a[0], a[1], b[0], b[1])
        pts = set(zip(rr.tolist(), cc.tolist()))
        assert a in pts and b in pts and val.min() >= 0 and val.max() <= 1.0 + 1e-6
        if a[0] == b[0] or a[1] == b[1]:
            on_line = [(r, c, v) for r, c, v in zip(rr, cc, val) if (r == a[0] if a[0] == b[0] else c == a[1])]
            assert all(abs(v - 1.0) < 1e-6 for _, _, v in on_line)


def _reference_modules():
    root = ref_shim.reference_root()
    if root is None:
        pytest.skip("reference tree not present")
    ref_shim.import_reference()
    import skimage.draw
    skimage.draw.line_aa = gto.line_aa            # scikit-image is absent: inject the restated algorithm
    import utils.data as rdata
    import utils.projection as rproj
    rdata.line_aa = gto.line_aa
    return rdata, rproj


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
def test_pieces_match_the_live_reference(preset):
    rdata, rproj = _reference_modules()
    pts2d, pts3d = gto.synthetic_keypoints(preset, 3, seed=1)
    for b in range(3):
        for v in range(2):
            ref = rproj.coord2d_to_heatmap(pts2d[b, v][1:], res=64, sigma=1.0)
            assert np.abs(gto.coord2d_to_heatmap(pts2d[b, v][1:]) - ref).max() < 1e-6
            hm, _, theta = rdata.get_limb_data(pts2d[b, v].copy(), pts3d[b, v], 64, 64, "line", sigma=1.0, joint_preset=preset)
            ohm, otheta = gto.limb_data(pts2d[b, v], pts3d[b, v], preset)
            assert np.abs(ohm - hm).max() < 1e-6 and np.abs(otheta - theta).max() < 1e-6


def _reference_lifting_input(preset, pts2d, pts3d, tmpdir):
    """the reference's own loader (process_frame_data) + the wrapper's concatenation, for one frame"""
    import torch
    from types import SimpleNamespace
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))
    import dataloader.data_loader as dl
    n = pts2d.shape[1]
    local = pts3d[0] - pts3d[0][:1] * 0           # any split of pts3d into local pose + pelvis reproduces pts3d
    frame = dict(gt_camera_2d_left=pts2d[0].astype(np.float64), gt_camera_2d_right=pts2d[1].astype(np.float64),
                 gt_local_pose=local.astype(np.float64), gt_pelvis_left=np.zeros(3), gt_pelvis_right=(pts3d[1][0] - pts3d[0][0]).astype(np.float64),
                 gt_local_rot=np.zeros((n, 3)), input_rgb_left=np.zeros((3, 8, 8), np.float32), input_rgb_right=np.zeros((3, 8, 8), np.float32))
    path = os.path.join(tmpdir, "frame.npy")
    np.save(path, frame, allow_pickle=True)
    dl.resize_rgb = lambda rgb, w, h: np.zeros((3, h, w), np.float32)
    J = n - 1
    opt = SimpleNamespace(load_size_heatmap=[64, 64], joint_preset=preset, stereo=True, num_heatmap=J, num_rot_heatmap=J,
                          estimate_head=preset == "UnrealEgo", model="egotap_autoencoder", heatmap_type="sin")
    d = dl.process_frame_data(path, opt)
    joint = torch.cat((d["gt_heatmap_left"], d["gt_heatmap_right"]), 0)          # egotap_autoencoder_model.py:180-181
    limb = torch.cat((d["gt_limb_heatmap_left"], d["gt_limb_heatmap_right"]), 0)
    return torch.cat((joint, limb), 0).numpy()                                     # :213


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
def test_lifting_input_matches_the_reference_loader(preset, tmp_path):
    _reference_modules()
    pts2d, pts3d = gto.synthetic_keypoints(preset, 2, seed=2)
    for b in range(2):
        # the loader forms the right view's 3-D points as local pose + right pelvis: build pts3d accordingly
        p3 = np.stack([pts3d[b, 0], pts3d[b, 0] + (pts3d[b, 1][0] - pts3d[b, 0][0])[None]])
        ref = _reference_lifting_input(preset, pts2d[b], p3, str(tmp_path))
        got = gto.lifting_input(pts2d[b, 0], pts2d[b, 1], p3[0], p3[1], preset)
        assert got.shape == ref.shape == (6 * (pts2d.shape[2] - 1), 64, 64)
        assert np.abs(got - ref).max() < 2e-6


def test_golden_written_from_the_live_reference():
    d = np.load(GOLD)
    for preset in ("UnrealEgo", "EgoCap"):
        pts2d, pts3d = d[preset + "_pts2d"], d[preset + "_pts3d"]
        got = np.stack([gto.lifting_input(pts2d[b, 0], pts2d[b, 1], pts3d[b, 0], pts3d[b, 1], preset) for b in range(pts2d.shape[0])])
        assert np.abs(got - d[preset + "_input"]).max() < 2e-6
