"""The ground-truth heatmap synthesis kernel (csrc/gt_heatmap.cu) against its oracle
(oracle/gt_heatmap_oracle.py, pinned to the live reference loader in tests/test_gt_heatmap_oracle.py) and against the
golden written from the reference -- on the CUDA-on-CPU emulation of the kernel's own source ("emu", CPU suite) and
through the C ABI on the B200 ("cuda", -m gpu)."""
import os
import sys

import numpy as np
import pytest
import torch

import gt_heatmap_oracle as gto
import op_oracle

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "cuda_emu"))
sys.path.insert(0, os.path.dirname(__file__))
import build_emu  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_gt_heatmaps.npz")


@pytest.fixture(scope="module", params=["emu", pytest.param("cuda", marks=[pytest.mark.gpu, pytest.mark.timeout(600, method="thread")])])
def backend(request):
    if request.param == "emu":
        return build_emu.make_backend()[0]
    from gpu_adapter import GpuOpAdapter
    return GpuOpAdapter()


def _run(backend, pts2d, pts3d_left, preset):
    B, J = pts2d.shape[0], pts2d.shape[2] - 1
    out = torch.full((B, 6 * J, 64, 64), float("nan"))
    backend.gt_heatmaps(torch.from_numpy(pts2d).contiguous(), torch.from_numpy(pts3d_left).contiguous(), B, preset, out)
    return out.numpy()


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
def test_kernel_matches_oracle(backend, preset):
    pts2d, pts3d = gto.synthetic_keypoints(preset, 3, seed=5)
    # edge cases: a joint exactly on / just off the canvas borders, a zero-length limb, an axis-aligned limb
    pts2d[0, 0, 1] = (-64.0, 10.0); pts2d[0, 0, 2] = (1023.9, 1087.9); pts2d[0, 1, 3] = (1024.0, 500.0)
    pts2d[1, 0, 4] = pts2d[1, 0, 2]; pts2d[1, 1, 5] = pts2d[1, 1, 3] + np.float32([0.0, 160.0])
    got = _run(backend, pts2d, pts3d[:, 0], preset)
    want = np.stack([gto.lifting_input(pts2d[b, 0], pts2d[b, 1], pts3d[b, 0], pts3d[b, 0], preset) for b in range(3)])
    assert not np.isnan(got).any()
    # 4e-6 of the range (values reach 2): a few ulp of atanf / cosf / sinf and FMA contraction between libm and the device
    assert np.abs(got - want).max() < 4e-6 * max(1.0, np.abs(want).max())
    J = pts2d.shape[2] - 1
    assert np.abs(got[0, J + 2]).max() == 0.0          # joint 3 of the right view at x == 1024 -> empty map (x < res rule)


def test_kernel_matches_reference_golden(backend):
    d = np.load(GOLD)
    for preset in ("UnrealEgo", "EgoCap"):
        got = _run(backend, d[preset + "_pts2d"], np.ascontiguousarray(d[preset + "_pts3d"][:, 0]), preset)
        assert np.abs(got - d[preset + "_input"]).max() < 4e-6


def test_feeds_the_lifting_net_input_contract():
    """shape / channel order are the lifting input's: the oracle's split_input recovers the joint and limb stacks"""
    import egotap_oracle as orc
    pts2d, pts3d = gto.synthetic_keypoints("UnrealEgo", 1, seed=9)
    x = torch.from_numpy(gto.lifting_input(pts2d[0, 0], pts2d[0, 1], pts3d[0, 0], pts3d[0, 0], "UnrealEgo"))[None]
    pos, rot = orc.split_input(x, "UnrealEgo")
    assert pos.shape == (1, 30, 64, 64) and rot.shape == (1, 30, 2, 64, 64)
    # limb (view 0, joint 3): cos and sin maps are the same raw map scaled by cos / sin of one angle
    c, s = rot[0, 3, 0], rot[0, 3, 1]
    m = c.abs() > 1e-4
    ratio = (s[m] / c[m])
    assert ratio.numel() > 0 and (ratio - ratio[0]).abs().max() < 1e-3 * max(1.0, ratio[0].abs().item())
