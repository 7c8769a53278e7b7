"""Host-side orchestration of the training step (egotap_b200/training.py) executed on the CPU against the op oracle
(oracle/op_oracle.py: a torch restatement of every C-ABI op with the same argument meaning), and compared with
torch.autograd on the restated train-mode forward + loss + AdamW (oracle/train_oracle.py, itself pinned to the
unmodified reference in tests/test_train_oracle.py).

What this proves without a GPU: which buffer feeds which op, every stride / group layout / split-K chunking, the
gradient routing of all 100+ parameters, BatchNorm running-stat updates, the loss and the optimiser arithmetic.
What it cannot prove: the CUDA kernels behind the ops -- those are compared op by op with the same oracle methods in
tests/test_train_gpu.py."""
import pytest
import torch

import op_oracle
import train_oracle as tro
import weights
from egotap_b200 import training
from egotap_b200.synthetic import synthetic_heatmaps


def _inputs(preset, batch):
    x = synthetic_heatmaps(preset, batch, seed=17, kind="gauss")
    g = torch.Generator().manual_seed(19)
    nj = 16 if preset == "UnrealEgo" else 17
    gt = torch.randn(batch, nj, 3, generator=g) * 20
    return x, gt


def _engine(preset, precision, exact=False, **kw):
    sd = weights.make_state_dict(preset, seed=5)
    params = {k: v.clone().contiguous() for k, v in sd.items()}
    eng = training.TrainEngine(preset, params, precision=precision, backend=op_oracle.OracleBackend(exact=exact), **kw)
    return sd, params, eng


def _rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


@pytest.mark.parametrize("preset,batch", [("UnrealEgo", 3), ("EgoCap", 2)])
def test_orchestration_matches_autograd_exactly(preset, batch):
    """exact=True stores every operand "pair" in fp32, so the only difference to autograd is summation order:
    every gradient, BatchNorm buffer and AdamW update must agree to fp32 rounding.  This is the check of the
    host-side dataflow (buffers, strides, groups, split-K, routing)."""
    sd, params, eng = _engine(preset, "bf16x3", exact=True, attn_chunk=2)   # attn_chunk < batch: chunk loop runs
    x, gt = _inputs(preset, batch)
    ref_loss, ref_sd, ref_state, ref_grads = tro.train_step(sd, x, gt, preset)
    pose = eng.forward(x.clone())
    ref_pose, _ = tro.forward_train(sd, x, preset)
    assert _rel(pose, ref_pose) < 2e-5
    loss = eng.loss_and_grad(gt.clone())
    assert abs(float(loss[0]) - float(ref_loss)) < 2e-6 * max(1.0, abs(float(ref_loss)))
    seen = []
    grads = eng.backward(on_stage=lambda i, a, b: seen.append((i, a, b)))
    assert [s[0] for s in seen] == list(range(len(eng.stages))) and seen[-1][2] == eng.flat_grad.numel()
    assert not torch.isnan(eng.flat_grad).any()
    for k, g_ref in ref_grads.items():
        if g_ref is None:
            continue
        # the absolute floor covers gradients that are analytically zero and pure rounding noise on both sides
        # (key biases: softmax shift invariance; Linear / LayerNorm biases in front of a train-mode BatchNorm)
        err, scale = (grads[k] - g_ref).abs().max().item(), g_ref.abs().max().item()
        assert err <= 2e-4 * scale + 1e-7, (k, err, scale)
    for k in sd:
        if "running_" in k:
            assert _rel(params[k], ref_sd[k]) < 1e-5, k
        if k.endswith("num_batches_tracked"):
            assert int(params[k]) == int(ref_sd[k])
    eng.adamw_step(lr=1e-3, eps=1e-4)       # compare the update, not the weight
    for k, g_ref in ref_grads.items():
        if g_ref is None:
            continue
        upd_ref, upd = ref_sd[k] - sd[k], params[k] - sd[k]
        assert (upd - upd_ref).abs().max().item() <= 2e-2 * upd_ref.abs().max().item() + 2e-7, k
    for k in sd:                            # dead parameters: no gradient -> AdamW skips them (reference behaviour)
        if "cls_token" in k or "pooler" in k:
            assert torch.equal(params[k], sd[k])


def test_bf16x3_mode_gradient_error():
    """the fp32-parity operand mode (bf16 hi/lo pairs): gradients downstream of every LeakyReLU (head, propagation
    chain) agree to 2e-3 of the tensor maximum; the FC encoders and the ViT are compared by direction and norm because a
    ~1e-5 forward difference can flip individual LeakyReLU'(z) factors at z ~ 0 in an FC block (60-90 BatchNorm rows
    here), which moves single rows of dW by O(10 %) of the maximum without being an error of either side"""
    preset, batch = "UnrealEgo", 3
    sd, params, eng = _engine(preset, "bf16x3")
    x, gt = _inputs(preset, batch)
    ref_loss, _, _, ref_grads = tro.train_step(sd, x, gt, preset)
    pose = eng.forward(x.clone())
    assert _rel(pose, tro.forward_train(sd, x, preset)[0]) < 5e-4
    assert abs(float(eng.loss_and_grad(gt.clone())[0]) - float(ref_loss)) < 2e-5 * max(1.0, abs(float(ref_loss)))
    grads = eng.backward()
    for k, g_ref in ref_grads.items():
        if g_ref is None or g_ref.abs().max() < 1e-7:
            continue
        upstream = "heatmap_encoder" in k          # anything at or upstream of a LeakyReLU (all six FC blocks, the ViT)
        if not upstream:
            err, scale = (grads[k] - g_ref).abs().max().item(), g_ref.abs().max().item()
            assert err <= 2e-3 * scale + 1e-7, (k, err, scale)
        assert _cos(grads[k], g_ref) > 0.9995, (k, _cos(grads[k], g_ref))
        assert abs(float(grads[k].norm() / g_ref.norm()) - 1) < 2e-2, k


def test_bf16_mode_gradients_are_close():
    """plain-bf16 operands (config 5's precision): gradients agree with fp32 autograd in direction, with the
    stated looser bound (cosine > 0.97 per tensor at batch 2, where BatchNorm sees 60 rows; for the tensors that carry most of the gradient mass)"""
    preset, batch = "UnrealEgo", 2
    sd, params, eng = _engine(preset, "bf16")
    x, gt = _inputs(preset, batch)
    _, _, _, ref_grads = tro.train_step(sd, x, gt, preset)
    eng.forward(x.clone())
    eng.loss_and_grad(gt.clone())
    grads = eng.backward()
    for k in ("pose_mlp.pose_fcs.0.weight", "skel_sequential_layer.lstm_custom.layers.0.h2h.weight",
              "pos_heatmap_encoder.fc1.fc.weight", "rot_heatmap_encoder.fc1.fc.weight",
              "pos_heatmap_encoder.vit.encoder.layer.2.intermediate.dense.weight",
              "pos_heatmap_encoder.vit.encoder.layer.0.attention.attention.value.weight",
              "pos_heatmap_encoder.vit.embeddings.patch_embeddings.projection.weight"):
        assert _cos(grads[k], ref_grads[k]) > 0.97, (k, _cos(grads[k], ref_grads[k]))


def test_second_step_repacks_and_moves_loss():
    """two full train_step calls: weights are re-packed after the optimiser step and the loss changes accordingly"""
    preset = "UnrealEgo"
    sd, params, eng = _engine(preset, "bf16x3")
    x, gt = _inputs(preset, 2)
    l0 = float(eng.train_step(x, gt)[0])
    l1 = float(eng.train_step(x, gt)[0])
    _, new_sd, state, _ = tro.train_step(sd, x, gt, preset)
    ref_l1 = float(tro.train_step(new_sd, x, gt, preset, opt_state=state)[0])
    assert l1 != l0
    assert abs(l1 - ref_l1) < 5e-4 * max(1.0, abs(ref_l1))


def test_splitk_chunks_stay_inside_padding():
    for rows in (1, 30, 64, 480, 3840, 7680, 18432, 147456, 147457):
        ld = training.pad_ld(rows)
        for n, k in ((1024, 256), (1024, 1024), (4096, 1024), (2048, 16384), (128, 512), (768, 256), (2048, 512)):
            G, chunk = training.splitk(n, k, rows)
            assert chunk % 64 == 0 and G >= 1 and G * chunk >= rows and G * chunk <= ld and G <= training.SPLITK_MAX


def test_grad_layout_covers_every_trained_parameter():
    for preset in ("UnrealEgo", "EgoCap"):
        sd = weights.make_state_dict(preset, seed=5)
        names = [k for _, ks in training.param_order(preset) for k in ks]
        trained = [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k
                   and "cls_token" not in k and "pooler" not in k]
        assert sorted(names) == sorted(trained)
