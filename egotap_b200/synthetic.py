"""Synthetic UnrealEgo/EgoCap-shaped lifting inputs (no dataset is reachable offline).

Shapes and value ranges follow what the reference's data pipeline feeds the lifting net
(reference ``dataloader/data_loader.py:127,193-199``, ``utils/projection.py:263-279``,
``utils/data.py:197-252``): per view J joint heatmaps (sigma = 1 px Gaussian, peak ~1) and
J limb heatmaps (2 x a sigma = 1 blurred parent->child line) modulated by cos / sin of the
limb's elevation angle (same angle in both views).  Channel layout of the ``(B, 6J, 64, 64)``
tensor: ``[joint L | joint R | cos L | sin L | cos R | sin R]``
(reference ``model/egotap_autoencoder_model.py:182-183,215``).
"""
import math

import torch

PRESET_J = {"UnrealEgo": 15, "EgoCap": 17}
HM = 64


def synthetic_heatmaps(preset="UnrealEgo", batch=16, seed=1234, kind="gauss", device="cpu"):
    """Return a contiguous fp32 ``(batch, 6J, 64, 64)`` tensor on ``device``.

    ``kind='gauss'`` is the realistic distribution above; ``kind='uniform'`` is U[0,1) stress input.
    The generator is seeded on the CPU so values are identical on every device."""
    J = PRESET_J[preset]
    g = torch.Generator().manual_seed(seed)
    if kind == "uniform":
        return torch.rand(batch, 6 * J, HM, HM, generator=g).to(device)
    if kind != "gauss":
        raise ValueError("kind must be 'gauss' or 'uniform', got %r" % (kind,))
    ys = torch.arange(HM, dtype=torch.float32).view(1, 1, HM, 1)
    xs = torch.arange(HM, dtype=torch.float32).view(1, 1, 1, HM)
    # joints: one integer-centred Gaussian per (view, joint)
    c = torch.randint(0, HM, (batch, 2 * J, 2), generator=g).float()
    cx, cy = c[..., 0, None, None], c[..., 1, None, None]
    joint = torch.exp(-0.5 * ((xs - cx) ** 2 + (ys - cy) ** 2))
    # limbs: distance-to-segment profile, 2 * (1/sqrt(2 pi)) peak, times cos/sin(theta)
    p = torch.rand(batch, 2 * J, 2, 2, generator=g) * (HM - 1)
    ax, ay = p[..., 0, 0, None, None], p[..., 0, 1, None, None]
    bx, by = p[..., 1, 0, None, None], p[..., 1, 1, None, None]
    dx, dy = bx - ax, by - ay
    t = (((xs - ax) * dx + (ys - ay) * dy) / (dx * dx + dy * dy + 1e-6)).clamp_(0, 1)
    d2 = (xs - (ax + t * dx)) ** 2 + (ys - (ay + t * dy)) ** 2
    limb = (2.0 / math.sqrt(2 * math.pi)) * torch.exp(-0.5 * d2)            # (B, 2J, 64, 64)
    theta = (torch.rand(batch, J, generator=g) - 0.5) * math.pi             # per limb, both views
    cos = torch.cos(theta)[:, None, :, None, None]
    sin = torch.sin(theta)[:, None, :, None, None]
    limb = limb.view(batch, 2, J, HM, HM)
    rot = torch.stack([limb * cos, limb * sin], dim=2)                      # (B, view, {cos,sin}, J, ..)
    x = torch.cat([joint, rot.reshape(batch, 4 * J, HM, HM)], dim=1).contiguous()
    return x.to(device)
