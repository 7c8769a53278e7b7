"""Drop-in at the reference's own seam, exercised with the REAL reference wrapper on CPU (skipped where the
reference tree is absent): `network.define_AutoEncoder` is swapped for `egotap_b200.define_AutoEncoder`
(INTEGRATION.md section 1), then the unmodified `create_model(opt)` -> `EgoTAPAutoEncoderModel.initialize`,
`load_networks('best')` (strict state_dict load of a checkpoint written by the reference's own lifting net),
`eval()` / `set_eval_mode()` run as in reference test.py:28-37.  The forward itself needs a GPU: on CPU the wrapper
must reach our module and get its explicit no-CPU-path error (never a silent fallback)."""
import contextlib
import io
import os

import pytest
import torch

import ref_shim

pytestmark = pytest.mark.skipif(ref_shim.reference_root() is None, reason="reference tree not present")


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
def test_reference_wrapper_accepts_the_swapped_factory(preset, tmp_path, state_dicts):
    na = ref_shim.import_reference()
    import model.network as ref_network
    from model.models import create_model

    import egotap_b200
    opt = ref_shim.make_opt(preset, isTrain=False, log_dir=str(tmp_path), experiment_name="exp", use_amp=False,
                            path_to_trained_heatmap=None, model_name="resnet18", init_ImageNet=False,
                            use_gt_heatmap=True, distributed=False)
    original = ref_network.define_AutoEncoder
    ref_network.define_AutoEncoder = egotap_b200.define_AutoEncoder          # <- the one-line swap
    try:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            model = create_model(opt)
    finally:
        ref_network.define_AutoEncoder = original
    assert isinstance(model.net_AutoEncoder, egotap_b200.EgoTAPAutoEncoder)
    assert model.model_names == ["HeatMap", "RotHeatMap", "AutoEncoder"]

    # a checkpoint written by the REFERENCE's lifting net (and the wrapper's own heatmap nets) ...
    save_dir = os.path.join(str(tmp_path), "exp")
    os.makedirs(save_dir)
    with contextlib.redirect_stdout(io.StringIO()):
        ref_net = na.EgoTAPAutoEncoder(opt, input_channel_scale=2)
    sd = state_dicts(preset)
    ref_net.load_state_dict(sd, strict=True)
    torch.save(ref_net.cpu().state_dict(), os.path.join(save_dir, "best_net_AutoEncoder.pth"))
    torch.save(model.net_HeatMap.state_dict(), os.path.join(save_dir, "best_net_HeatMap.pth"))
    torch.save(model.net_RotHeatMap.state_dict(), os.path.join(save_dir, "best_net_RotHeatMap.pth"))
    # ... loads through the reference's loader (strict) into our module
    with contextlib.redirect_stdout(io.StringIO()):
        model.load_networks("best")
    for k, v in model.net_AutoEncoder.state_dict().items():
        assert torch.equal(v, sd[k]), k
    model.eval()
    model.set_eval_mode()
    assert not model.net_AutoEncoder.training
    # the wrapper's forward reaches our module; without a GPU that is an explicit error, not a fallback
    J = 15 if preset == "UnrealEgo" else 17
    model.gt_heatmap_left = model.gt_heatmap_right = torch.zeros(1, J, 64, 64)
    model.gt_limb_heatmap_left = model.gt_limb_heatmap_right = torch.zeros(1, 2 * J, 64, 64)
    model.input_rgb_left = model.input_rgb_right = torch.zeros(1, 3, 256, 256)
    with pytest.raises(RuntimeError, match="no CPU path"):
        model.forward(evaluate=True)
    assert model.pred_heatmap_cat.shape == (1, 6 * J, 64, 64)      # the wrapper built our input in its own layout
