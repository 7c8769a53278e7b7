"""GPU test of the opt-in CUDA-graph replay of the inference forward (opt ``b200_cuda_graph``; small-batch serving).
Reference eval scripts run batch 16 / 32 (scripts/test/unrealego.sh); this is the serving path for such batches."""
import pytest
import torch

import egotap_oracle as orc
from ref_shim import make_opt

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
def test_graph_replay_equals_eager_forward(preset, state_dicts):
    """same kernels, same buffers: the replayed forward is bit-identical to the launched one, for several batch sizes, after a
    weight update (packed copies are refreshed outside the graph) and for batches above the graph limit (eager path)"""
    import egotap_b200
    from egotap_b200 import synthetic_heatmaps
    sd = state_dicts(preset)
    nets = []
    for graph in (0, 4):
        net = egotap_b200.EgoTAPAutoEncoder(make_opt(preset, b200_cuda_graph=graph), input_channel_scale=2)
        net.load_state_dict(sd, strict=True)
        nets.append(net.cuda().eval())
    eager, graphed = nets
    def same(batch, seed):
        x = synthetic_heatmaps(preset, batch, seed=seed, kind="gauss").cuda()
        a, b = eager.predict_pose(x), graphed.predict_pose(x)
        assert torch.equal(a, b), (batch, (a - b).abs().max().item())
        return x, b
    for batch, seed in ((1, 3), (4, 4), (1, 5)):
        same(batch, seed)
    assert sorted(graphed._graphs) == [1, 4] and graphed._plan_batch == 4      # plan sized for the graph limit up front
    # above the graph limit: launched, not replayed.  The plan grows (geometrically), its workspace is re-created, so the captured
    # graphs are dropped and re-captured lazily by the next small batch
    same(6, 6)
    assert graphed._graphs == {} and graphed._plan_batch == 8
    same(4, 7)
    x, b = same(3, 8)
    assert sorted(graphed._graphs) == [3, 4]
    with torch.no_grad():
        ref = orc.forward(sd, x.cpu()[:2], preset)
    assert orc.parity_report(b[:2], ref)["rel"] <= 5e-4
    with torch.no_grad():
        for net in nets:
            net.pose_mlp.pose_fcs._modules["0"].bias.add_(0.5)
    x = synthetic_heatmaps(preset, 4, seed=9, kind="gauss").cuda()
    a, b = eager.predict_pose(x), graphed.predict_pose(x)
    assert torch.equal(a, b)
