"""egotap_b200 -- B200-native (sm_100a) implementation of EgoTAP's heatmap->3D lifting path.

Public surface mirrors the reference's construction seam:
``define_AutoEncoder(opt, model)`` (reference model/network.py:24-33) and the
``EgoTAPAutoEncoder`` nn.Module (reference model/net_architecture.py:579-758).
"""
from .net_architecture import EgoTAPAutoEncoder  # noqa: F401
from .network import define_AutoEncoder  # noqa: F401
from .options import make_opt  # noqa: F401
from .synthetic import synthetic_heatmaps  # noqa: F401

__all__ = ["EgoTAPAutoEncoder", "define_AutoEncoder", "make_opt", "synthetic_heatmaps"]
