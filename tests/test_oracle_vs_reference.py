"""Live check of the oracle against the real reference tree (skipped where it is absent,
e.g. on the GPU box).  CPU only."""
import pytest
import torch

import egotap_oracle as orc
import ref_shim
from egotap_b200.synthetic import synthetic_heatmaps

pytestmark = pytest.mark.skipif(ref_shim.reference_root() is None, reason="reference tree not present")


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
def test_live_reference(preset, state_dicts):
    net = ref_shim.build_reference_net(preset)
    sd = state_dicts(preset)
    assert set(net.state_dict()) == set(sd)
    net.load_state_dict(sd, strict=True)
    x = synthetic_heatmaps(preset, 2, seed=99, kind="gauss")
    with torch.no_grad():
        ref = net(x)[0]
        mine = orc.forward(sd, x, preset)
        truth = orc.forward({k: v.double() for k, v in sd.items()}, x.double(), preset)
    assert orc.parity_report(mine, ref)["rel"] < 2e-5
    # the fp64 run of the oracle is the tighter truth both fp32 runs sit next to
    assert orc.parity_report(ref, truth)["rel"] < 2e-5
    assert orc.parity_report(mine, truth)["rel"] < 2e-5
