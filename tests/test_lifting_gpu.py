"""GPU parity tests proper: the CUDA path (through the C ABI, via the reference-shaped nn.Module)
against the CPU oracle on identical seeded inputs and weights, stage by stage and end to end, and
against the committed outputs of the unmodified reference (tests/golden).

Tolerances (BASELINE.json north_star): fp32-parity mode (bf16x3 operands, fp32 accumulate):
max|d| / max|ref| <= 1e-3 and MPJPE delta <= 0.1 mm -- asserted 2x tighter (5e-4) end to end and on every
intermediate stage, 10x tighter on MPJPE (measured: pose 2.1e-4, ViT stages 2e-5); bf16-operand mode carries its own looser, stated bound
(rel <= 5e-2, MPJPE delta <= 0.5 mm; PyTorch's own CPU bf16 autocast of the reference lands at
1.4e-2..3.3e-2 / 0.20..0.24 mm).
"""
import json
import os

import numpy as np
import pytest
import torch

import egotap_oracle as orc
from ref_shim import make_opt

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
OUT = os.path.join(os.path.dirname(os.path.dirname(__file__)), "gpurun_out")


def _record(name, rep):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, "parity.jsonl")
    with open(path, "a") as f:
        f.write(json.dumps(dict(test=name, **rep)) + "\n")


def _module(preset, precision, sd):
    import egotap_b200
    net = egotap_b200.EgoTAPAutoEncoder(make_opt(preset, b200_precision=precision), input_channel_scale=2)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval()


def _token_perm(preset):
    """heatmap-major token -> raster token of the 24x24 mosaic (reference net_architecture.py:397-402)."""
    g = orc.geometry(preset)
    perm = []
    for n in range(g["grid"] ** 2):
        for pr in range(4):
            for pc in range(4):
                perm.append(((n // g["grid"]) * 4 + pr) * g["side"] + (n % g["grid"]) * 4 + pc)
    return torch.tensor(perm)


def _rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
@pytest.mark.parametrize("precision,tol", [("bf16x3", 5e-4), ("bf16", 5e-2)])
def test_stages_match_oracle(preset, precision, tol, state_dicts):
    from egotap_b200 import capi, synthetic_heatmaps
    sd = state_dicts(preset)
    net = _module(preset, precision, sd)
    x = synthetic_heatmaps(preset, 2, seed=1234, kind="gauss")
    taps = {}
    with torch.no_grad():
        ref_pose = orc.forward(sd, x, preset, taps=taps)
    g = orc.geometry(preset)
    perm = _token_perm(preset)
    xc = x.cuda()
    B, J, live = 2, g["J"], g["n_hm"] * 16
    rep = {}
    stages = [("embeddings", 0), ("layer0", 1), ("layer1", 2), ("layer2", 3)]
    for name, stage in stages:
        net._run(xc, last_stage=stage)
        torch.cuda.synchronize()
        hid = net._debug_buffer("hidden", (B, 576, 1024)).clone()
        # after the LAST layer only the live tokens are defined: the dummy/mask-token rows are attended to as
        # keys/values but never read again, so the row-wise GEMMs of the last layer skip them (UnrealEgo geometry)
        keep = live if name == "layer2" else 576
        rep[name] = _rel(hid[:, :keep], taps[name][:, perm][:, :keep])
    net._run(xc, last_stage=4)
    fin = net._debug_buffer("fin_hi", (B, live, 1024), torch.bfloat16).float()
    if precision == "bf16x3":
        fin = fin + net._debug_buffer("fin_lo", (B, live, 1024), torch.bfloat16).float()
    rep["vit_out"] = _rel(fin, taps["vit_out"][:, perm][:, :live])
    net._run(xc, last_stage=6)
    emb = net._debug_buffer("embed", (B, J, 512)).clone()
    rep["pos_embed"] = _rel(emb[..., :256], taps["pos_embed"])
    rep["rot_embed"] = _rel(emb[..., 256:], taps["rot_embed"])
    net._run(xc, last_stage=7)
    rep["skel"] = _rel(net._debug_buffer("skel", (B, J, 512)).clone(), taps["skel"])
    pose = net(xc)[0]
    torch.cuda.synchronize()
    rep.update({"pose_" + k: v for k, v in orc.parity_report(pose, ref_pose).items()})
    _record("stages_%s_%s" % (preset, precision), rep)
    unscaled = ("pose_max_abs_mm", "pose_mpjpe_delta_mm", "pose_per_joint_rel", "pose_joints_counted", "pose_joints_total")
    bad = {k: v for k, v in rep.items() if k not in unscaled and not (v <= tol)}
    assert not bad, (bad, rep)
    assert rep["pose_mpjpe_delta_mm"] <= (0.01 if precision == "bf16x3" else 0.5), rep
    # BASELINE.json north_star: "per-joint 3D coordinates within 1e-3 relative" (fp32-accumulate mode); the bf16-operand mode
    # has its own stated bound: 5e-2 of the batch maximum (tol above) and 1e-1 per joint
    assert rep["pose_per_joint_rel"] <= (1e-3 if precision == "bf16x3" else 1e-1), rep
    assert rep["pose_joints_counted"] >= 0.9 * rep["pose_joints_total"], rep


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
@pytest.mark.parametrize("kind", ["gauss", "uniform"])
def test_matches_reference_golden(preset, kind, state_dicts):
    """fp32-parity mode vs outputs of the unmodified reference module (tests/golden/make_golden.py)."""
    from egotap_b200 import synthetic_heatmaps
    gold = np.load(os.path.join(GOLD, "ref_%s_%s.npz" % (preset, kind)))
    _, iseed, batch = (int(v) for v in gold["meta"])
    net = _module(preset, "bf16x3", state_dicts(preset))
    x = synthetic_heatmaps(preset, batch, seed=iseed, kind=kind).cuda()
    pose, rot, indep, hm = net(x)
    rep = orc.parity_report(pose, torch.from_numpy(gold["pose"]))
    _record("golden_%s_%s" % (preset, kind), rep)
    assert rep["rel"] <= 5e-4 and rep["mpjpe_delta_mm"] <= 0.01 and rep["per_joint_rel"] <= 1e-3, rep
    # quirks that are specification: 4-tuple of the reference's shapes, aux outputs all zero, head joint last
    assert [rot.shape[1], indep.shape[1], hm.shape[1]] == list(gold["aux_shapes"]) and hm.shape == x.shape
    assert rot.abs().max() == 0 and indep.abs().max() == 0 and hm.abs().max() == 0
    assert torch.equal(net.predict_pose(x), pose)


def test_batch_sizes_and_ragged_batches(state_dicts):
    """Frames are independent: any batch split gives the same rows (incl. batch 1 and a non-multiple of the tile)."""
    from egotap_b200 import synthetic_heatmaps
    preset = "UnrealEgo"
    net = _module(preset, "bf16x3", state_dicts(preset))
    x = synthetic_heatmaps(preset, 7, seed=3, kind="gauss").cuda()
    full = net.predict_pose(x).clone()
    one = net.predict_pose(x[3:4]).clone()
    part = net.predict_pose(x[:5]).clone()
    # not bit-identical: at <= 32 frames the first FC block is split along its reduction dimension into a batch-dependent number
    # of chunks (fp32 summation order), which shows at the 1e-5 level of the fp32-parity mode (vs the oracle: ~1e-4)
    scale = full.abs().max().item()
    assert (full[3:4] - one).abs().max().item() < 1e-4 * scale       # measured 3e-5 (round 2, several boxes)
    assert (full[:5] - part).abs().max().item() < 1e-4 * scale
    with torch.no_grad():
        ref = orc.forward(state_dicts(preset), x.cpu(), preset)
    assert orc.parity_report(full, ref)["rel"] <= 5e-4


@pytest.mark.parametrize("preset,precision,tol", [("UnrealEgo", "bf16x3", 5e-4), ("EgoCap", "bf16", 5e-2)])
def test_full_size_batch_equals_small_batches(preset, precision, tol, state_dicts):
    """BASELINE.json's full sizes through a size-independent property: frames are independent, so a batch of 1100
    (more than one 1024-frame chunk of the persistent chain kernel, several waves of every GEMM) must reproduce,
    row for row, what the same frames give in batches of 256 / 76 -- which the oracle tests pin at small size.
    Cases: config 2's preset / precision, and config 3's (EgoCap, bf16 operands, batch >= 1024)."""
    from egotap_b200 import synthetic_heatmaps
    net = _module(preset, precision, state_dicts(preset))
    base = synthetic_heatmaps(preset, 44, seed=21, kind="gauss").cuda()
    x = base.repeat(25, 1, 1, 1) * torch.linspace(0.5, 1.5, 1100, device="cuda")[:, None, None, None]
    big = net.predict_pose(x).clone()
    parts = torch.cat([net.predict_pose(x[i:i + 256]).clone() for i in range(0, 1024, 256)] + [net.predict_pose(x[1024:]).clone()])
    assert torch.isfinite(big).all()
    assert (big - parts).abs().max().item() < 2e-5 * max(1.0, big.abs().max().item())
    with torch.no_grad():
        ref = orc.forward(state_dicts(preset), x[1084:].cpu(), preset)
    rep = orc.parity_report(big[1084:], ref)                      # 16 frames of the last (ragged) chunk vs the oracle
    _record("full_size_%s_%s" % (preset, precision), rep)
    assert rep["rel"] <= tol, rep


def test_per_joint_launch_path_agrees_with_persistent_chain(state_dicts):
    """The persistent propagation-chain kernel vs the per-joint launch sequence it replaced (both CUDA)."""
    import subprocess, sys, os, json
    code = ("import sys; sys.path[:0]=[%r,%r]; import torch, weights, egotap_b200; from ref_shim import make_opt;"
            "net=egotap_b200.EgoTAPAutoEncoder(make_opt('EgoCap'),2); net.load_state_dict(weights.make_state_dict('EgoCap',5));"
            "net=net.cuda().eval(); x=egotap_b200.synthetic_heatmaps('EgoCap',5,seed=8).cuda();"
            "print(net.predict_pose(x).flatten().tolist())")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    # (the third run folds the six in-layer LayerNorm kernels into the GEMMs around them: opt-in EGOTAP_LN=fold)
    for env in ({}, {"EGOTAP_PU": "steps", "EGOTAP_ATTN": "unfused", "EGOTAP_SKIP_DUMMY": "0"}, {"EGOTAP_LN": "fold"}):
        r = subprocess.run([sys.executable, "-c", code % (root, os.path.join(root, "oracle"))], capture_output=True,
                           text=True, env=dict(os.environ, **env), timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(torch.tensor(json.loads(r.stdout.strip().splitlines()[-1])))
    assert ((outs[0] - outs[1]).abs().max() / outs[1].abs().max()).item() < 2e-4
    assert ((outs[0] - outs[2]).abs().max() / outs[2].abs().max()).item() < 2e-4


@pytest.mark.parametrize("precision", ["bf16x3", "bf16"])
def test_forward_is_bit_deterministic(precision, state_dicts):
    """No kernel of the path uses atomics or an order-dependent reduction, so repeating a forward must give the same bits; a rare
    protocol race in a warp-specialised kernel (the first two-issuer attention build had one, found on hardware) shows up as a
    repetition that differs.  tools/soak.py is the long version (profiles/r02i_soak.txt: 2,400 forwards, 2,400 attention launches)."""
    from egotap_b200 import synthetic_heatmaps
    preset = "UnrealEgo"
    net = _module(preset, precision, state_dicts(preset))
    for batch in (16, 96):
        x = synthetic_heatmaps(preset, batch, seed=batch, kind="gauss").cuda()
        first = net.predict_pose(x).clone()
        for _ in range(25):
            assert torch.equal(net.predict_pose(x), first)


def test_empty_ragged_and_odd_inputs(state_dicts):
    """Edge cases the reference accepts: empty batch, fp16 / fp64 heatmaps, a non-contiguous view, a second CUDA
    stream; all-zero heatmaps (an undetected person) must give finite poses."""
    from egotap_b200 import synthetic_heatmaps
    preset = "UnrealEgo"
    sd = state_dicts(preset)
    net = _module(preset, "bf16x3", sd)
    x = synthetic_heatmaps(preset, 3, seed=31, kind="gauss").cuda()
    base = net.predict_pose(x).clone()
    out = net(x[:0])
    assert out[0].shape == (0, 16, 3) and out[3].shape == (0, 90, 64, 64)
    assert (net.predict_pose(x.double()) - base).abs().max() < 1e-6
    half = net.predict_pose(x.half())
    with torch.no_grad():
        ref_half = orc.forward(sd, x.half().float().cpu(), preset)
    assert orc.parity_report(half, ref_half)["rel"] <= 5e-4
    wide = torch.zeros(3, 90, 64, 128, device="cuda")
    wide[..., ::2] = x
    assert (net.predict_pose(wide[..., ::2]) - base).abs().max() < 1e-6
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        on_side = net.predict_pose(x).clone()
    side.synchronize()
    assert (on_side - base).abs().max() < 1e-6
    zeros = net.predict_pose(torch.zeros_like(x))
    with torch.no_grad():
        ref0 = orc.forward(sd, torch.zeros(1, 90, 64, 64), preset)
    assert torch.isfinite(zeros).all() and orc.parity_report(zeros[:1], ref0)["rel"] <= 5e-4


def test_cpu_input_is_an_error_not_a_fallback(state_dicts):
    preset = "EgoCap"
    net = _module(preset, "bf16x3", state_dicts(preset))
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 102, 64, 64))
    with pytest.raises(AssertionError):
        net(torch.zeros(1, 90, 64, 64, device="cuda"))


def test_weight_update_repacks(state_dicts):
    """In-place parameter updates (an optimizer step, load_state_dict) must invalidate the packed copies."""
    from egotap_b200 import synthetic_heatmaps
    preset = "UnrealEgo"
    net = _module(preset, "bf16x3", state_dicts(preset))
    x = synthetic_heatmaps(preset, 1, seed=11, kind="gauss").cuda()
    a = net.predict_pose(x).clone()
    with torch.no_grad():
        net.pose_mlp.pose_fcs._modules["0"].bias.add_(1.0)
    b = net.predict_pose(x).clone()
    assert torch.allclose(b[:, :15], a[:, :15] + 1.0, atol=1e-5)
    net.load_state_dict(state_dicts(preset))
    assert torch.allclose(net.predict_pose(x), a, atol=1e-6)
