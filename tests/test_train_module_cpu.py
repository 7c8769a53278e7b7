"""The train-mode seam of the drop-in module and the data-parallel gradient exchange, on the CPU: the module's
autograd Function and the staged all-reduce are exercised with the op oracle as the engine backend (the product
backend is the CUDA library; tests/test_zz_train_gpu.py repeats this on the B200)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import op_oracle
from oracle_backed import use_oracle_backend
import train_oracle as tro
import weights
from ref_shim import make_opt
from egotap_b200.synthetic import synthetic_heatmaps


def _module(preset, sd):
    import egotap_b200
    net = egotap_b200.EgoTAPAutoEncoder(make_opt(preset), input_channel_scale=2)
    net.load_state_dict(sd, strict=True)
    use_oracle_backend(net)
    return net


def test_reference_style_training_loop_on_the_module():
    """net.train(); loss in torch; loss.backward(); torch.optim.AdamW.step() -- the reference's optimize_parameters
    (model/egotap_autoencoder_model.py:299-323) -- against the training oracle"""
    preset = "UnrealEgo"
    sd = weights.make_state_dict(preset, seed=5)
    net = _module(preset, sd)
    x = synthetic_heatmaps(preset, 2, seed=17, kind="gauss")
    gt = torch.randn(2, 16, 3, generator=torch.Generator().manual_seed(19)) * 20
    ref_loss, ref_sd, _, ref_grads = tro.train_step(sd, x, gt, preset)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, eps=1e-4, weight_decay=0.0)
    net.train()
    opt.zero_grad()
    pose, rot, indep, hm = net(x)
    assert pose.requires_grad and pose.shape == (2, 16, 3) and rot.shape == (2, 45) and hm.shape == x.shape
    loss = tro.total_loss(pose, gt, preset)
    assert abs(float(loss) - float(ref_loss)) < 1e-5
    loss.backward()
    named = dict(net.named_parameters())
    for k, g in ref_grads.items():
        if g is None:
            assert named[k].grad is None or float(named[k].grad.abs().max()) == 0.0, k
            continue
        assert (named[k].grad - g).abs().max().item() <= 2e-4 * g.abs().max().item() + 1e-7, k
    assert named["pos_heatmap_encoder.vit.embeddings.cls_token"].grad is None
    opt.step()
    new_sd = net.state_dict()
    for k, v in ref_sd.items():
        if not v.is_floating_point():
            assert int(new_sd[k]) == int(v), k
            continue
        upd_ref, upd = v - sd[k], new_sd[k] - sd[k]
        assert (upd - upd_ref).abs().max().item() <= 2e-2 * upd_ref.abs().max().item() + 2e-7, k
    # second iteration: the engine re-packs the updated weights and gradients are fresh tensors (no aliasing of .grad)
    g_before = named["pose_mlp.pose_fcs.0.weight"].grad.clone()
    opt.zero_grad(set_to_none=False)
    tro.total_loss(net(x)[0], gt, preset).backward()
    assert not torch.equal(named["pose_mlp.pose_fcs.0.weight"].grad, g_before)
    # eval mode still has no CPU path
    with pytest.raises(RuntimeError, match="no CPU path"):
        net.eval()(x)


def test_train_mode_without_grad_is_a_plain_forward():
    preset = "EgoCap"
    sd = weights.make_state_dict(preset, seed=5)
    net = _module(preset, sd).train()
    x = synthetic_heatmaps(preset, 2, seed=3, kind="gauss")
    with torch.no_grad():
        pose = net.predict_pose(x)
    ref, _ = tro.forward_train(sd, x, preset)
    assert not pose.requires_grad and (pose - ref).abs().max() < 2e-5 * ref.abs().max()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle")):
            if p not in sys.path:
                sys.path.insert(0, p)
        from egotap_b200 import ddp, training
        torch.set_num_threads(2)
        preset = "UnrealEgo"
        sd = weights.make_state_dict(preset, seed=5)
        params = {k: v.clone().contiguous() for k, v in sd.items()}
        eng = training.TrainEngine(preset, params, precision="bf16x3", backend=op_oracle.OracleBackend(exact=True))
        red = ddp.StagedGradAllReduce(eng, bucket_elems=4 * 1024 * 1024)
        x = synthetic_heatmaps(preset, 1, seed=100 + rank, kind="gauss")
        gt = torch.randn(1, 16, 3, generator=torch.Generator().manual_seed(200 + rank)) * 20
        eng.train_step(x, gt, reducer=red)
        # numpy copies: torch tensors travel through multiprocessing queues as shared-memory handles, which die with
        # the worker
        probe = {k: params[k].numpy().copy() for k in ("pose_mlp.pose_fcs.0.weight",
                                                "pos_heatmap_encoder.vit.encoder.layer.0.attention.attention.query.weight",
                                                "pos_heatmap_encoder.fc1.bn.running_mean")}
        q.put((rank, probe, red.launched, eng.flat_grad.numel()))
    finally:
        dist.destroy_process_group()


def test_staged_gradient_allreduce_over_gloo():
    """world_size 2: every rank trains on its own micro-batch; after the staged SUM all-reduce + AdamW(grad_scale=1/2)
    the trained weights are identical on both ranks and equal to AdamW on the AVERAGE of the two ranks' gradients;
    BatchNorm running buffers stay per rank (no SyncBN in the reference)"""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=900) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # buckets tile the flat buffer exactly, in order
    launched, total = got[0][2], got[0][3]
    assert launched[0][0] == 0 and launched[-1][1] == total and all(a[1] == b[0] for a, b in zip(launched, launched[1:]))
    assert len(launched) >= 3
    w0, w1 = ({k: torch.from_numpy(v) for k, v in g[1].items()} for g in got)
    for k in ("pose_mlp.pose_fcs.0.weight", "pos_heatmap_encoder.vit.encoder.layer.0.attention.attention.query.weight"):
        assert torch.equal(w0[k], w1[k]), k
    assert not torch.equal(w0["pos_heatmap_encoder.fc1.bn.running_mean"], w1["pos_heatmap_encoder.fc1.bn.running_mean"])
    # expected: AdamW on the mean of the two per-rank autograd gradients
    preset = "UnrealEgo"
    sd = weights.make_state_dict(preset, seed=5)
    grads = []
    for rank in range(world):
        x = synthetic_heatmaps(preset, 1, seed=100 + rank, kind="gauss")
        gt = torch.randn(1, 16, 3, generator=torch.Generator().manual_seed(200 + rank)) * 20
        grads.append(tro.train_step(sd, x, gt, preset)[3])
    for k in ("pose_mlp.pose_fcs.0.weight", "pos_heatmap_encoder.vit.encoder.layer.0.attention.attention.query.weight"):
        g = (grads[0][k] + grads[1][k]) / 2
        want, _, _ = tro.adamw_update(sd[k], g, torch.zeros_like(g), torch.zeros_like(g), 1, lr=1e-3, eps=1e-4)
        upd_ref, upd = want - sd[k], w0[k] - sd[k]
        assert (upd - upd_ref).abs().max().item() <= 2e-2 * upd_ref.abs().max().item() + 2e-7, k
