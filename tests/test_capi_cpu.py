"""CPU-side checks of the boundary: the C-ABI library loads without a GPU and exports every symbol
include/egotap_b200.h declares; host-side argument errors are reported, not fatal; the Python module
keeps the reference's surface (no compute calls here)."""
import os
import re

import pytest
import torch

from ref_shim import make_opt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from egotap_b200.build import build
    build()
    from egotap_b200 import capi
    return capi.lib()


def test_library_exports_every_declared_symbol(lib):
    from egotap_b200 import capi
    header = open(os.path.join(ROOT, "include", "egotap_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(egotap_b200_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "libegotap_b200.so does not export %s" % name
    assert sorted(capi.EXPORTS) == declared
    assert lib.egotap_b200_abi_version() == 1


def test_param_order_is_the_reference_key_set(lib):
    import weights
    from egotap_b200 import capi
    for preset, n_total in (("UnrealEgo", 117), ("EgoCap", 115)):
        pid = capi.PRESET_ID[preset]
        names = [lib.egotap_b200_param_name(pid, i).decode() for i in range(lib.egotap_b200_num_params(pid))]
        keys = [k for k, _, _ in weights.key_shapes(preset)]
        assert len(keys) == n_total and set(names) <= set(keys)
        dead = set(keys) - set(names)
        assert all(("cls_token" in k) or ("pooler" in k) or k.endswith("num_batches_tracked") for k in dead), dead


def test_plan_sizes_and_argument_errors(lib):
    import ctypes as C
    pb, wb = C.c_size_t(), C.c_size_t()
    assert lib.egotap_b200_plan_sizes(0, 0, 256, C.byref(pb), C.byref(wb)) == 0
    assert 350e6 < pb.value < 450e6          # ~97 M parameters as bf16 hi+lo
    assert 4e9 < wb.value < 16e9
    pb1 = C.c_size_t()
    assert lib.egotap_b200_plan_sizes(0, 1, 256, C.byref(pb1), C.byref(wb)) == 0 and pb1.value < pb.value
    assert lib.egotap_b200_plan_sizes(7, 0, 256, C.byref(pb), C.byref(wb)) < 0
    assert b"preset" in lib.egotap_b200_last_error()
    assert lib.egotap_b200_plan_sizes(0, 0, 0, C.byref(pb), C.byref(wb)) < 0


def test_module_surface_matches_reference():
    import egotap_b200
    import weights
    for preset in ("UnrealEgo", "EgoCap"):
        net = egotap_b200.EgoTAPAutoEncoder(make_opt(preset), input_channel_scale=2)
        sd = weights.make_state_dict(preset, seed=1)
        assert list(net.state_dict().keys()) == [k for k, _, _ in weights.key_shapes(preset)]
        net.load_state_dict(sd, strict=True)
        for k, v in net.state_dict().items():
            assert v.dtype == sd[k].dtype and tuple(v.shape) == tuple(sd[k].shape)
        with pytest.raises(RuntimeError, match="no CPU path"):
            net.eval()(torch.zeros(1, net.channels_heatmap, 64, 64))
    with pytest.raises(ValueError):
        egotap_b200.EgoTAPAutoEncoder(make_opt("UnrealEgo", joint_preset="Nope"), input_channel_scale=2)
    with pytest.raises(NotImplementedError):
        egotap_b200.EgoTAPAutoEncoder(make_opt("UnrealEgo", skel_layer="LSTM"), input_channel_scale=2)
    with pytest.raises(Exception):
        egotap_b200.define_AutoEncoder(make_opt("UnrealEgo"), "something_else")


def test_product_options_helper_matches_the_reference_presets():
    """egotap_b200.options.make_opt (product side) == the fields the reference parser yields for the two presets
    (ref_shim.make_opt is the test-side twin used to construct the real reference module)."""
    import egotap_b200
    for preset in ("UnrealEgo", "EgoCap"):
        a, b = vars(egotap_b200.make_opt(preset)), vars(make_opt(preset))
        for k in b:
            assert a[k] == b[k], (preset, k)
    with pytest.raises(ValueError):
        egotap_b200.make_opt("Nope")


def test_factory_inits_like_the_reference(capsys):
    import egotap_b200
    net = egotap_b200.define_AutoEncoder(make_opt("EgoCap"), "egotap_autoencoder")
    out = capsys.readouterr().out
    assert "total number of parameters of AutoEncoder" in out and "96.938 M" in out
    sd = net.state_dict()
    w = sd["pos_heatmap_encoder.fc1.fc.weight"]
    assert abs(w.std().item() - (2.0 / 16384) ** 0.5) < 2e-4                 # kaiming fan_in
    assert sd["pos_heatmap_encoder.fc1.fc.bias"].abs().max() == 0
    assert sd["pos_heatmap_encoder.fc1.bn.running_var"].min() == 1
    assert abs(sd["pos_heatmap_encoder.vit.embeddings.position_embeddings"].std().item() - 0.02) < 1e-3   # trunc-normal(0.02)
