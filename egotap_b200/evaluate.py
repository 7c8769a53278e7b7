"""On-GPU evaluation metrics with the reference's surface.

``compute_metrics(pred_pose, gt_pose, running_average_dict)`` mirrors reference
``utils/evaluate.py:54-73``: per-frame MPJPE and PA-MPJPE in millimetres (poses are in cm, ``cm2mm = 10``,
reference ``model/egotap_autoencoder_model.py:100``), accumulated into the caller's running-average dict.
One fused kernel computes both metrics for the whole batch (3x3 Procrustes SVD in registers) and there is a
single device->host copy per batch instead of the reference's per-frame synchronisations."""
import torch

from . import capi

cm2mm = 10


def pose_metrics(pred_pose, gt_pose, unit_scale=cm2mm):
    """(B, J, 3) CUDA tensors -> (mpjpe, pa_mpjpe), each (B,) fp32 on the GPU."""
    capi.require_cuda(pred_pose, gt_pose)
    if pred_pose.shape != gt_pose.shape or pred_pose.dim() != 3 or pred_pose.shape[-1] != 3:
        raise ValueError("expected matching (B, J, 3) poses, got %s and %s" % (tuple(pred_pose.shape), tuple(gt_pose.shape)))
    p = pred_pose.detach().float().contiguous()
    g = gt_pose.detach().float().contiguous()
    B, J = p.shape[0], p.shape[1]
    out = torch.empty((2, B), dtype=torch.float32, device=p.device)
    capi.check(capi.lib().egotap_b200_pose_metrics(p.data_ptr(), g.data_ptr(), B, J, float(unit_scale), out[0].data_ptr(),
                                                   out[1].data_ptr(), capi.current_stream()), "pose_metrics")
    return out[0], out[1]


def compute_metrics(pred_pose, gt_pose, running_average_dict):
    """Reference-compatible: returns CPU tensors (mpjpes, pa_mpjpes) and updates ``running_average_dict`` per frame."""
    m, pa = pose_metrics(pred_pose, gt_pose)
    both = torch.stack([m, pa]).cpu()          # the one device->host copy of the batch
    mpjpes, pa_mpjpes = both[0], both[1]
    for i in range(mpjpes.shape[0]):
        running_average_dict.update(dict(mpjpe=mpjpes[i], pa_mpjpe=pa_mpjpes[i]))
    return mpjpes, pa_mpjpes
