"""Ground-truth heatmap synthesis on the GPU.

With ``--use_gt_heatmap`` the reference builds the lifting network's input on the CPU for every frame -- Gaussian joint
heatmaps (reference utils/projection.py:263-279), anti-aliased limb lines blurred and modulated by the cosine / sine of the
limb's elevation angle (utils/data.py:175-252, dataloader/data_loader.py:127-132,193-199) -- and ships 1.47 MB per frame to
the device.  ``synthesize`` produces the same (B, 6J, 64, 64) tensor from the keypoints directly in HBM with one kernel
launch, so only ~0.5 KB per frame crosses PCIe and the result is already where the lifting kernels read it.
No CPU fallback: CPU tensors raise."""
import torch

from . import capi

_POINTS = {"UnrealEgo": 16, "EgoCap": 18}


def synthesize(pts2d, pts3d_left, preset="UnrealEgo", out=None):
    """pts2d: (B, 2, J+1, 2) keypoints of the left / right view in 1024-pixel image coordinates (``gt_camera_2d_left``,
    ``gt_camera_2d_right`` of the reference's frame files); pts3d_left: (B, J+1, 3) = ``gt_local_pose + gt_pelvis_left``.
    Returns (B, 6J, 64, 64) fp32 = [joint L | joint R | cos L | sin L | cos R | sin R] on the same device."""
    if preset not in _POINTS:
        raise ValueError("joint_preset is {} which is undefined".format(preset))
    capi.require_cuda(pts2d, pts3d_left)
    n = _POINTS[preset]
    B = pts2d.shape[0]
    if tuple(pts2d.shape) != (B, 2, n, 2) or tuple(pts3d_left.shape) != (B, n, 3):
        raise ValueError("expected pts2d (B, 2, %d, 2) and pts3d_left (B, %d, 3), got %s and %s"
                         % (n, n, tuple(pts2d.shape), tuple(pts3d_left.shape)))
    pts2d, pts3d_left = pts2d.float().contiguous(), pts3d_left.float().contiguous()
    if out is None:
        out = torch.empty((B, 6 * (n - 1), 64, 64), dtype=torch.float32, device=pts2d.device)
    if B == 0:
        return out
    with torch.cuda.device(pts2d.device):
        capi.check(capi.lib().egotap_b200_gt_heatmaps(pts2d.data_ptr(), pts3d_left.data_ptr(), B, capi.PRESET_ID[preset],
                                                      out.data_ptr(), capi.current_stream()), "gt_heatmaps")
    return out
