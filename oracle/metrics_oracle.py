"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the evaluation metrics.

Restates reference ``utils/util.py:328-379`` (batch_compute_similarity_transform_torch) and the per-frame MPJPE /
PA-MPJPE of ``utils/evaluate.py:54-73`` in fp64.  Pinned against the reference's own function, run live in the
build container (tests/test_metrics_oracle.py) and through committed outputs (tests/golden/ref_metrics.npz)."""
import torch


def procrustes_align(pred, gt):
    """pred, gt: (B, J, 3) -> pred aligned to gt by the optimal similarity transform (scale, rotation, translation)."""
    S1, S2 = pred.double().transpose(1, 2), gt.double().transpose(1, 2)        # (B, 3, J)
    mu1, mu2 = S1.mean(-1, keepdim=True), S2.mean(-1, keepdim=True)
    X1, X2 = S1 - mu1, S2 - mu2
    var1 = (X1 ** 2).sum(dim=(1, 2))
    K = X1 @ X2.transpose(1, 2)
    U, s, Vh = torch.linalg.svd(K)
    V = Vh.transpose(1, 2)
    Z = torch.eye(3, dtype=torch.float64).repeat(K.shape[0], 1, 1)
    Z[:, -1, -1] *= torch.sign(torch.det(U @ V.transpose(1, 2)))
    R = V @ Z @ U.transpose(1, 2)
    scale = torch.diagonal(R @ K, dim1=1, dim2=2).sum(-1) / var1
    t = mu2 - scale[:, None, None] * (R @ mu1)
    return (scale[:, None, None] * (R @ S1) + t).transpose(1, 2)


def pose_metrics(pred, gt, unit_scale=10.0):
    """per-frame (mpjpe, pa_mpjpe), fp64, in units * unit_scale (cm -> mm)."""
    mpjpe = (gt.double() - pred.double()).norm(dim=-1).mean(-1) * unit_scale
    pa = (gt.double() - procrustes_align(pred, gt)).norm(dim=-1).mean(-1) * unit_scale
    return mpjpe, pa
