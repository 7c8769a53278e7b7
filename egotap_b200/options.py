"""The ``opt`` fields the lifting net reads, as the reference's option parser would produce them.

The reference configures everything through one flat argparse namespace (``options/*.py``); the lifting net reads
``joint_preset, ae_hidden_size, heatmap_type, num_heatmap, num_rot_heatmap, estimate_head, patched_heatmap_ae,
skel_layer, load_size_heatmap, stereo, gpu_ids, init_type`` (reference ``model/net_architecture.py:585-662``,
``model/network.py:24-33``).  ``make_opt`` builds that namespace for the two dataset presets with the flag values of
``scripts/test/unrealego.sh`` / ``egocap.sh`` and the preset rules of ``options/dataset_options.py:30-41`` -- for
callers (benchmarks, serving code) that do not go through the reference's command line."""
from types import SimpleNamespace

_JOINTS = {"UnrealEgo": 15, "EgoCap": 17}


def make_opt(joint_preset="UnrealEgo", **overrides):
    if joint_preset not in _JOINTS:
        raise ValueError("joint_preset is {} which is undefined".format(joint_preset))
    ue = joint_preset == "UnrealEgo"
    opt = dict(joint_preset=joint_preset, model="egotap_autoencoder", ae_hidden_size=128, heatmap_type="sin",
               num_heatmap=_JOINTS[joint_preset], num_rot_heatmap=_JOINTS[joint_preset], estimate_head=ue, stereo=True,
               patched_heatmap_ae=True, skel_layer="PU", load_size_heatmap=[64, 64], gpu_ids=[], init_type="kaiming",
               model_name="resnet18", init_ImageNet=False)
    opt.update(overrides)
    return SimpleNamespace(**opt)
