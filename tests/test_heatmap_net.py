"""(f1) heatmap-producer mirror vs the reference module (CPU, live; skipped where the reference is absent) and the
hand-off layout.  The producer is a torch module by design; the lifting net behind it is covered elsewhere."""
import contextlib
import io

import pytest
import torch

import ref_shim
from ref_shim import make_opt


@pytest.mark.skipif(ref_shim.reference_root() is None, reason="reference tree not present")
@pytest.mark.parametrize("which", ["pos", "rot"])
def test_mirror_matches_reference_module(which):
    na = ref_shim.import_reference()
    from egotap_b200.heatmap_net import HeatMapUNet
    opt = make_opt("UnrealEgo", init_ImageNet=False, model_name="resnet18")
    if which == "pos":
        opt.num_rot_heatmap = 0
    else:
        opt.num_heatmap = 0
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        ref = na.HeatMap_UnrealEgo_Shared(opt, "resnet18", input_channel_scale=2).eval()
    torch.manual_seed(0)
    mine = HeatMapUNet(opt, "resnet18", input_channel_scale=2).eval()
    assert list(mine.state_dict().keys()) == list(ref.state_dict().keys())
    ref.load_state_dict(mine.state_dict(), strict=True)
    g = torch.Generator().manual_seed(1)
    l, r = torch.rand(1, 3, 256, 256, generator=g), torch.rand(1, 3, 256, 256, generator=g)
    with torch.no_grad():
        a, b = mine(l, r), ref(l, r)
    assert a.shape == b.shape == (1, 30 if which == "pos" else 60, 64, 64)
    assert (a - b).abs().max() <= 1e-5 * b.abs().max()


def test_handoff_slots_alias_the_lifting_input():
    import egotap_b200
    from egotap_b200.heatmap_net import StereoPoseEstimator
    opt = make_opt("EgoCap")
    lift = egotap_b200.EgoTAPAutoEncoder(opt, input_channel_scale=2)
    est = StereoPoseEstimator(None, None, lift)
    buf, pos, rot = est.heatmap_buffer(3, torch.device("cpu"))
    assert buf.shape == (3, 102, 64, 64) and pos.shape[1] == 34 and rot.shape[1] == 68
    pos.fill_(1.0); rot.fill_(2.0)
    assert buf[:, :34].eq(1).all() and buf[:, 34:].eq(2).all()          # views of ONE buffer, reference channel order
    assert buf.data_ptr() == pos.data_ptr()


@pytest.mark.gpu
def test_end_to_end_rgb_to_pose_matches_staged_computation(state_dicts):
    """config 4 data flow: producers (torch) -> in-place hand-off -> lifting kernels == the same producers' fp32
    heatmaps concatenated the reference way and lifted by the oracle."""
    import egotap_b200
    import egotap_oracle as orc
    from egotap_b200.heatmap_net import HeatMapUNet, StereoPoseEstimator
    preset = "UnrealEgo"
    sd = state_dicts(preset)
    lift = egotap_b200.EgoTAPAutoEncoder(make_opt(preset), input_channel_scale=2)
    lift.load_state_dict(sd)
    pos_opt, rot_opt = make_opt(preset), make_opt(preset)
    pos_opt.num_rot_heatmap = 0
    rot_opt.num_heatmap = 0
    torch.manual_seed(0)
    hp, hr = HeatMapUNet(pos_opt), HeatMapUNet(rot_opt)
    est = StereoPoseEstimator(hp, hr, lift, producer_dtype=None).cuda().eval()
    g = torch.Generator().manual_seed(2)
    l, r = torch.rand(2, 3, 256, 256, generator=g), torch.rand(2, 3, 256, 256, generator=g)
    pose = est(l.cuda(), r.cuda())
    with torch.no_grad():
        hm = torch.cat([hp.cpu()(l, r), hr.cpu()(l, r)], dim=1)           # reference forward_heatmap layout
        ref = orc.forward(sd, hm, preset)
    rep = orc.parity_report(pose, ref)
    assert rep["rel"] <= 1e-3 and rep["mpjpe_delta_mm"] <= 0.1, rep      # includes cuDNN-vs-CPU conv differences


@pytest.mark.gpu
def test_pipeline_replayed_from_a_cuda_graph_equals_the_launched_one(state_dicts):
    """StereoPoseEstimator(cuda_graph=True): both producers, the hand-off and the lifting kernels captured once per batch
    shape and replayed -- same kernels on the same buffers, so the poses are bit-identical to the launched pipeline, also
    for a second batch (inputs are copied into the graph's static buffers) and a second batch size"""
    import egotap_b200
    from egotap_b200.heatmap_net import HeatMapUNet, StereoPoseEstimator
    preset = "EgoCap"
    lift = egotap_b200.EgoTAPAutoEncoder(make_opt(preset), input_channel_scale=2)
    lift.load_state_dict(state_dicts(preset))
    pos_opt, rot_opt = make_opt(preset), make_opt(preset)
    pos_opt.num_rot_heatmap = 0
    rot_opt.num_heatmap = 0
    torch.manual_seed(0)
    hp, hr = HeatMapUNet(pos_opt), HeatMapUNet(rot_opt)
    eager = StereoPoseEstimator(hp, hr, lift).cuda().eval()
    graphed = StereoPoseEstimator(hp, hr, lift, cuda_graph=True).cuda().eval()
    g = torch.Generator().manual_seed(3)
    for batch in (2, 2, 3, 2):          # batch 3 grows the lifting plan: the batch-2 graph is stale and must be re-captured
        l, r = torch.rand(batch, 3, 256, 256, generator=g).cuda(), torch.rand(batch, 3, 256, 256, generator=g).cuda()
        a = eager(l, r).clone()
        b = graphed(l, r).clone()
        assert torch.isfinite(a).all() and torch.equal(a, b), (batch, (a - b).abs().max().item())
    assert len(graphed._graphs) == 2
