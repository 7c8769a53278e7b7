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


def save_predictions(pred_poses, gt_poses, input_paths, save_path, data_dir):
    """The result files of reference ``utils/evaluate.py:127-144`` (``test_evaluate(save_result=True)``), byte for byte:
    ``pred_pose.npy`` (N, num_joints, 3) fp32 in ``save_path``; ``gt_<dataset>_pose.npy`` and ``input_<dataset>_paths.npy``
    one directory up; ``input_paths.pkl``.  ``pred_poses`` / ``gt_poses``: per-batch arrays or tensors (cm);
    ``input_paths``: per-batch sequences of frame paths."""
    import os
    import pickle

    import numpy as np

    def host(a):
        return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    pred = np.concatenate([host(p) for p in pred_poses], axis=0)
    gt = np.concatenate([host(g) for g in gt_poses], axis=0)
    paths = np.concatenate([np.asarray(p) for p in input_paths], axis=0).reshape(-1, 1)
    name = os.path.normpath(data_dir).split("/")[-1].lower()
    np.save(os.path.join(save_path, "pred_pose.npy"), pred)
    np.save(os.path.join(save_path, os.pardir, "gt_{}_pose.npy".format(name)), gt)
    np.save(os.path.join(save_path, os.pardir, "input_{}_paths.npy".format(name)), paths)
    with open(os.path.join(save_path, "input_paths.pkl"), "wb") as f:
        pickle.dump(paths, f)
    return pred, gt, paths
