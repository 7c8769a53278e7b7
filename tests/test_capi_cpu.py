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


def _header_prototypes():
    """name -> list of C parameter type strings, parsed from include/egotap_b200.h"""
    header = open(os.path.join(ROOT, "include", "egotap_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|long long|const char\*)\s+(egotap_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if args == ["void"] else args
    return protos


def _ctype_class(decl):
    """coarse class of a C parameter declaration: 'ptr', 'int', 'll', 'size', 'float'"""
    if "*" in decl:
        return "ptr"
    for key, cls in (("size_t", "size"), ("long long", "ll"), ("double", "double"), ("float", "float"), ("int", "int")):
        if decl.startswith(key) or (" " + key + " ") in (" " + decl):
            return cls
    raise AssertionError("unparsed parameter declaration %r" % decl)


def test_ctypes_signatures_match_the_header(lib):
    """every entry that has ctypes argtypes: same arity and the same coarse type class per position as the header --
    a swapped or missing argument in the binding would otherwise only show up on the GPU"""
    import ctypes as C
    protos = _header_prototypes()
    cls_of = {C.c_void_p: "ptr", C.c_char_p: "ptr", C.c_int: "int", C.c_longlong: "ll", C.c_size_t: "size", C.c_float: "float", C.c_double: "double"}
    checked = 0
    for name, params in protos.items():
        argtypes = getattr(lib, name).argtypes
        if argtypes is None:
            continue
        assert len(argtypes) == len(params), (name, len(argtypes), len(params))
        for i, (t, decl) in enumerate(zip(argtypes, params)):
            got = cls_of.get(t, "ptr")           # POINTER(...) types
            want = _ctype_class(decl)
            if want == "size" and got == "ll":
                continue
            assert got == want, (name, i, decl, t)
        checked += 1
    assert checked >= 40


def test_training_entries_reject_bad_arguments_without_a_gpu(lib):
    """argument validation runs before any CUDA call, so the marshalling order of the training entries can be probed on
    the CPU: a deliberately bad value in ONE position must produce the error that names it"""
    import ctypes as C
    buf = (C.c_float * 64)()
    p = C.cast(buf, C.c_void_p)
    err = lambda: lib.egotap_b200_last_error().decode()
    assert lib.egotap_b200_transpose_split(p, 128, 65, 128, 0, 0, p, p, 128, None, None, 0, 0, None, None, None, 0, None) < 0 and "cols (65)" in err()
    assert lib.egotap_b200_transpose_split(p, 128, 64, 64, 0, 0, None, None, 0, p, p, 130, 100, None, None, None, 0, None) < 0 \
        and "rows 128 pad 100 ld 130" in err()
    assert lib.egotap_b200_transpose_split(p, 128, 64, 64, 0, 0, p, p, 64, None, None, 0, 0, None, p, p, 100, None) < 0 \
        and "2 x 64 floats needed" in err()
    assert lib.egotap_b200_transpose_bf16(p, p, 10, 70, 128, 1, 0, 1, 0, p, p, 64, 0, 0, 64, None) < 0 and "cols (70)" in err()
    assert lib.egotap_b200_transpose_bf16(p, p, 10, 64, 128, 1, 0, 1, 0, p, p, 20, 0, 0, 8, None) < 0 and "rows 10 pad 8 ld 20" in err()
    assert lib.egotap_b200_colsum(p, 16, 6, 8, 0, 0, p, p, 64, None) < 0 and "multiples of 4" in err()
    assert lib.egotap_b200_reduce_partials(p, 2, 6, p, None) < 0 and "G 2 n 6" in err()
    assert lib.egotap_b200_softmax_bwd(p, p, 8, 512, 1.0, p, p, p, p, None) < 0 and "got 512" in err()
    assert lib.egotap_b200_layernorm_bwd(p, p, p, 1, 576, 600, 1e-12, p, 0, p, p, p, 64, None) < 0 and "rows_out" in err()
    assert lib.egotap_b200_bn_apply(p, 31, 128, p, p, p, p, 128, None, 0, 15, 0, None) < 0 and "rows (31)" in err()
    assert lib.egotap_b200_pu_cell_fwd(p, 2048, 2048, p, 768, 768, p, p, None, None, p, p, 15, 15, 4, None) < 0 and "step 15 of 15" in err()
    assert lib.egotap_b200_pu_cell_bwd(p, 2048, 2048, p, 768, 768, p, p, p, p, p, p, 2048, 2048, p, 768, 770, p, p, 0, 15, 4, None) < 0 \
        and "strides" in err()
    parents = (C.c_int * 16)(*[0, 0, 1, 1, 2, 3, 4, 5, 2, 3, 8, 9, 10, 11, 12, 13])
    assert lib.egotap_b200_pose_loss(p, p, 4, 16, parents, 15, 0, 0.1, -0.01, p, p, p, 64, None) < 0 and "16 joints / 15 parents" in err()
    assert lib.egotap_b200_head_bwd(p, p, 512, p, p, p, 4, 15, p, 512, p, p, p, None, None, p, 64, None) < 0 and "dWg" in err()
    assert lib.egotap_b200_embed_grads(p, 5, 30, p, p, None) < 0 and "geometry" in err()
    assert lib.egotap_b200_regroup_gather(p, 512, 2, 4, 15, 128, p, None) < 0 and "multiples of 4" in err()


def test_training_and_synthesis_have_no_cpu_path():
    """the product never falls back: without a CUDA device the training engine and the heatmap synthesis raise"""
    import weights
    from egotap_b200 import training
    from egotap_b200.gt_heatmaps import synthesize
    if torch.cuda.is_available():
        pytest.skip("needs a GPU-less host")
    sd = weights.make_state_dict("UnrealEgo", seed=1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        training.TrainEngine("UnrealEgo", {k: v.clone() for k, v in sd.items()})
    with pytest.raises(RuntimeError, match="no CPU path"):
        synthesize(torch.zeros(1, 2, 16, 2), torch.zeros(1, 16, 3), "UnrealEgo")
    import egotap_b200
    net = egotap_b200.EgoTAPAutoEncoder(make_opt("UnrealEgo"), input_channel_scale=2).train()
    with pytest.raises(RuntimeError, match="no CPU path"):
        net(torch.zeros(2, net.channels_heatmap, 64, 64))
    # nothing under egotap_b200/ imports the oracle
    pkg = os.path.join(ROOT, "egotap_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(import|from)\s+(op_oracle|egotap_oracle|train_oracle|gt_heatmap_oracle|metrics_oracle)", src, re.M), f
