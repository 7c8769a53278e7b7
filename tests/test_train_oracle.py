"""(f2 groundwork) the training-step oracle vs the unmodified reference module in train mode + the reference's
loss classes + torch.optim.AdamW, live on CPU (skipped where the reference is absent) and through a committed
golden.  No CUDA training path exists yet; this pins the checker the next round builds against."""
import os

import numpy as np
import pytest
import torch

import ref_shim
import train_oracle as tro
import weights
from egotap_b200.synthetic import synthetic_heatmaps

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_train_step.npz")
PROBE = ["pose_mlp.pose_fcs.0.weight", "pos_heatmap_encoder.fc3.fc.weight",
         "skel_sequential_layer.lstm_custom.layers.0.h2h.weight",
         "pos_heatmap_encoder.vit.encoder.layer.2.attention.attention.query.bias",
         "pos_heatmap_encoder.vit.embeddings.mask_token", "pos_heatmap_encoder.fc1.bn.running_var"]


def _inputs(preset, batch=3):
    x = synthetic_heatmaps(preset, batch, seed=17, kind="gauss")
    g = torch.Generator().manual_seed(19)
    nj = 16 if preset == "UnrealEgo" else 17
    gt = torch.randn(batch, nj, 3, generator=g) * 20
    return x, gt


def reference_train_step(preset):
    """The reference's own optimisation step (model/egotap_autoencoder_model.py:299-323 without AMP)."""
    ref_shim.import_reference()
    from utils.loss import LossFuncCosSim, LossFuncMPJPE
    net = ref_shim.build_reference_net(preset)
    sd = weights.make_state_dict(preset, seed=5)
    net.load_state_dict(sd, strict=True)
    net.train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, eps=1e-4, weight_decay=0.0)
    x, gt = _inputs(preset)
    opt.zero_grad()
    pose = net(x)[0]
    loss = LossFuncMPJPE()(pose, gt) * 0.1 + LossFuncCosSim(joint_preset=preset, estimate_head=preset == "UnrealEgo")(pose, gt) * -0.01 * 0.1
    loss.backward()
    opt.step()
    return float(loss), {k: v.detach().clone() for k, v in net.state_dict().items()}


@pytest.mark.skipif(ref_shim.reference_root() is None, reason="reference tree not present")
@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
def test_train_step_matches_live_reference(preset):
    ref_loss, ref_sd = reference_train_step(preset)
    sd = weights.make_state_dict(preset, seed=5)
    x, gt = _inputs(preset)
    loss, new_sd, state, grads = tro.train_step(sd, x, gt, preset)
    assert abs(float(loss) - ref_loss) < 1e-5 * max(1.0, abs(ref_loss))
    for k, v in ref_sd.items():
        if not v.is_floating_point():
            assert int(new_sd[k]) == int(v), k
            continue
        # one AdamW step moves every trained weight by ~lr; compare the UPDATE, not the weight
        upd_ref, upd = v - sd[k], new_sd[k] - sd[k]
        scale = max(upd_ref.abs().max().item(), 1e-12)
        # (key biases get an analytically zero gradient -- softmax is shift-invariant -- so their "updates" are
        # pure rounding noise around 1e-8: hence the absolute floor)
        assert (upd - upd_ref).abs().max().item() <= 2e-2 * scale + 2e-7, k


def test_train_step_matches_reference_golden():
    d = np.load(GOLD)
    preset = "UnrealEgo"
    sd = weights.make_state_dict(preset, seed=5)
    x, gt = _inputs(preset)
    loss, new_sd, _, _ = tro.train_step(sd, x, gt, preset)
    assert abs(float(loss) - float(d["loss"])) < 1e-5
    for i, k in enumerate(PROBE):
        ref_upd = torch.from_numpy(d["upd_%d" % i])
        upd = (new_sd[k] - sd[k]).flatten()[:ref_upd.numel()]
        assert (upd - ref_upd).abs().max() <= 2e-2 * ref_upd.abs().max() + 1e-9, k


def test_schedule_and_loss_pieces():
    assert tro.cosine_warmup_lr(0, 1e-3, 10, 100) == 0.0
    assert abs(tro.cosine_warmup_lr(10, 1e-3, 10, 100) - 1e-3) < 1e-12
    assert abs(tro.cosine_warmup_lr(55, 1e-3, 10, 100) - 0.5e-3) < 1e-9
    gt = torch.randn(4, 16, 3)
    assert abs(float(tro.loss_cos_sim(gt, gt, "UnrealEgo")) - 15.0) < 1e-4        # 15 bones, cos = 1 each
    assert float(tro.loss_mpjpe(gt, gt)) == 0.0
