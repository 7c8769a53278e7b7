"""TEST INFRASTRUCTURE ONLY -- deterministic state_dict generator with the reference's key set.

The reference has no seeding and its RNG stream is not something to depend on (SURVEY.md
section 8(a) a13), so parity tests never compare initialisations: they generate ONE state_dict
here and load it into the reference module, the oracle and the CUDA module alike.

Distributions follow the reference's random-init semantics
(``model/network_utils.py:37-58`` kaiming-normal fan_in on every Linear/Conv weight;
``model/modeling_vit.py:458-482`` trunc-normal(0.02) position embeddings / cls token;
BatchNorm1d / LayerNorm defaults), and -- with ``randomize=True`` -- additionally perturb
every term a fused/folded kernel could silently drop: biases, BN running stats and affine,
LN affine and the mask token.
"""
import math

import torch

from egotap_oracle import PRESETS, HID, EMB, PU_H


def key_shapes(preset="UnrealEgo", vit_layers=3):
    """(key, shape, kind) for every entry of net_AutoEncoder.state_dict()."""
    J = PRESETS[preset]["J"]
    out = []
    v = "pos_heatmap_encoder.vit."
    out += [(v + "embeddings.cls_token", (1, 1, HID), "tn"),
            (v + "embeddings.mask_token", (1, 1, HID), "mask"),
            (v + "embeddings.position_embeddings", (1, 576, HID), "tn"),
            (v + "embeddings.patch_embeddings.projection.weight", (HID, 1, 16, 16), "w"),
            (v + "embeddings.patch_embeddings.projection.bias", (HID,), "b")]
    for i in range(vit_layers):
        lp = v + "encoder.layer.%d." % i
        for n in ("query", "key", "value"):
            out += [(lp + "attention.attention.%s.weight" % n, (HID, HID), "w"),
                    (lp + "attention.attention.%s.bias" % n, (HID,), "b")]
        out += [(lp + "attention.output.dense.weight", (HID, HID), "w"),
                (lp + "attention.output.dense.bias", (HID,), "b"),
                (lp + "intermediate.dense.weight", (4 * HID, HID), "w"),
                (lp + "intermediate.dense.bias", (4 * HID,), "b"),
                (lp + "output.dense.weight", (HID, 4 * HID), "w"),
                (lp + "output.dense.bias", (HID,), "b"),
                (lp + "layernorm_before.weight", (HID,), "ln_w"),
                (lp + "layernorm_before.bias", (HID,), "ln_b"),
                (lp + "layernorm_after.weight", (HID,), "ln_w"),
                (lp + "layernorm_after.bias", (HID,), "ln_b")]
    out += [(v + "layernorm.weight", (HID,), "ln_w"), (v + "layernorm.bias", (HID,), "ln_b"),
            (v + "pooler.dense.weight", (HID, HID), "w"), (v + "pooler.dense.bias", (HID,), "b")]
    for enc, k1 in (("pos_heatmap_encoder", 16 * HID), ("rot_heatmap_encoder", 2 * 64 * 64)):
        for name, (n, k) in (("fc1", (2048, k1)), ("fc2", (512, 2048)), ("fc3", (EMB, 512))):
            p = "%s.%s." % (enc, name)
            out += [(p + "fc.weight", (n, k), "w"), (p + "fc.bias", (n,), "b"),
                    (p + "bn.weight", (n,), "bn_w"), (p + "bn.bias", (n,), "bn_b"),
                    (p + "bn.running_mean", (n,), "bn_m"), (p + "bn.running_var", (n,), "bn_v"),
                    (p + "bn.num_batches_tracked", (), "count")]
    s = "skel_sequential_layer.lstm_custom.layers."
    X = 2 * EMB
    for name, (n, k) in (("0.x2f", (PU_H + X, X)), ("0.x2h", (4 * PU_H, X)), ("0.b2h", (4 * PU_H, X)),
                         ("0.h2h", (4 * PU_H, PU_H)), ("1.x2f", (PU_H, PU_H)), ("1.x2h", (4 * PU_H, PU_H)),
                         ("1.h2h", (4 * PU_H, PU_H))):
        out += [(s + name + ".weight", (n, k), "w"), (s + name + ".bias", (n,), "b")]
    out += [("pose_mlp.pose_fcs.0.weight", (3, X + PU_H), "w"), ("pose_mlp.pose_fcs.0.bias", (3,), "b")]
    if PRESETS[preset]["global_head"]:
        out += [("global_mlp.pose_fcs.0.weight", (6, J * PU_H), "w"), ("global_mlp.pose_fcs.0.bias", (6,), "b")]
    return out


def make_state_dict(preset="UnrealEgo", seed=0, randomize=True, vit_layers=3):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape, kind in key_shapes(preset, vit_layers):
        if kind == "w":
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif kind == "tn":
            t = (torch.randn(shape, generator=g) * 0.02).clamp_(-0.04, 0.04)
        elif kind == "count":
            t = torch.tensor(0, dtype=torch.int64)
        elif not randomize:
            t = torch.ones(shape) if kind in ("ln_w", "bn_w", "bn_v") else torch.zeros(shape)
        elif kind == "b":
            t = torch.randn(shape, generator=g) * 0.05
        elif kind == "mask":
            t = torch.randn(shape, generator=g) * 0.5
        elif kind in ("ln_w", "bn_w"):
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif kind in ("ln_b", "bn_b"):
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_m":
            t = torch.randn(shape, generator=g) * 0.3
        elif kind == "bn_v":
            t = 0.5 + 1.5 * torch.rand(shape, generator=g)
        else:
            raise AssertionError(kind)
        sd[key] = t
    return sd
