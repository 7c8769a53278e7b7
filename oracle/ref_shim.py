"""TEST INFRASTRUCTURE ONLY -- import shim for the *real* EgoTAP reference tree.

Locates the read-only reference checkout (``$EGOTAP_REF`` or ``/root/reference``) and
makes ``model.net_architecture`` importable under the container's newer library
versions.  Nothing under ``egotap_b200/`` may import this file; it is used by
``tests/`` (skipped when the reference is absent, e.g. on the GPU box) and by
``tests/golden/make_golden.py`` to pin the clean-room oracle in ``oracle/egotap_oracle.py``.

The shims touch only non-hot-path imports:
  * ``transformers.pytorch_utils.find_pruneable_heads_and_indices`` (removed in v5; used only by
    ``prune_heads``, reference ``model/modeling_vit.py:36,284-300``, never called)
  * ``PreTrainedModel.get_head_mask`` (removed in v5; ``modeling_vit.py:590`` calls it with
    ``head_mask=None`` -> list of ``None``)
  * ``matplotlib`` / ``mpl_toolkits`` / ``skimage`` / ``natsort`` (plotting / data prep only)
"""
import os
import sys
import types
from types import SimpleNamespace

REF_CANDIDATES = [os.environ.get("EGOTAP_REF", ""), "/root/reference",
                  os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")]


def reference_root():
    for p in REF_CANDIDATES:
        if p and os.path.isfile(os.path.join(p, "model", "net_architecture.py")):
            return p
    return None


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Return the reference's ``model.net_architecture`` module (raises if the tree is absent)."""
    root = reference_root()
    if root is None:
        raise ImportError("EgoTAP reference tree not found (set $EGOTAP_REF)")
    sys.dont_write_bytecode = True
    import transformers.pytorch_utils as pu
    if not hasattr(pu, "find_pruneable_heads_and_indices"):
        pu.find_pruneable_heads_and_indices = lambda *a, **k: (set(), None)
    from transformers import PreTrainedModel
    if not hasattr(PreTrainedModel, "get_head_mask"):
        PreTrainedModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
        mpl.use = lambda *a, **k: None
        _stub("mpl_toolkits")
        _stub("mpl_toolkits.mplot3d", Axes3D=object)
    try:
        import skimage  # noqa: F401
    except Exception:
        sk = _stub("skimage")
        sk.draw = _stub("skimage.draw", line_aa=None)
    try:
        import natsort  # noqa: F401
    except Exception:
        _stub("natsort", natsorted=sorted)
    if root not in sys.path:
        sys.path.insert(0, root)
    import model.net_architecture as na  # noqa: E402
    return na


def make_opt(preset="UnrealEgo", **over):
    """The ``opt`` fields the lifting net reads (reference ``net_architecture.py:585-662``,
    flag values from ``scripts/test/unrealego.sh`` / ``egocap.sh``)."""
    ue = preset == "UnrealEgo"
    d = dict(joint_preset=preset, ae_hidden_size=128, heatmap_type="sin",
             num_heatmap=15 if ue else 17, num_rot_heatmap=15 if ue else 17,
             estimate_head=ue, patched_heatmap_ae=True, skel_layer="PU",
             load_size_heatmap=[64, 64], stereo=True, gpu_ids=[], init_type="kaiming",
             model="egotap_autoencoder")
    d.update(over)
    return SimpleNamespace(**d)


def build_reference_net(preset="UnrealEgo"):
    na = import_reference()
    opt = make_opt(preset)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        net = na.EgoTAPAutoEncoder(opt, input_channel_scale=2)
    return net.eval()
