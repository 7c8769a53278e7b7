import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
sys.dont_write_bytecode = True


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


# GPU files in the order they are run: the inference path BASELINE.json names first, then the section-8(f) widenings.  Every GPU
# test is an ordinary strict test (round 1's non-strict first-hardware-run quarantine is gone: all of it passed on the B200).
_GPU_FILE_ORDER = ["test_gemm_gpu.py", "test_ops_gpu.py", "test_lifting_gpu.py", "test_metrics.py", "test_heatmap_net.py",
                   "test_gt_heatmaps.py", "test_train_kernels.py", "test_zz_train_gpu.py", "test_zzz_graph_inference_gpu.py",
                   "test_zzz_epilogue_coalesced_gpu.py"]


# CPU tests that run in a background child process started with their module (tests/test_tensorcore_emu.py): collected last
# within the module so that the wait for the child overlaps with the module's in-process tests
_COLLECT_LAST_IN_MODULE = {"test_whole_training_step_on_product_kernel_source", "test_engine_with_persistent_bptt_matches_per_joint_path",
                           "test_whole_inference_path_on_product_source"}


def pytest_collection_modifyitems(config, items):
    module_pos = {}
    for item in items:
        module_pos.setdefault(str(item.fspath), len(module_pos))

    def key(item):
        name = os.path.basename(str(item.fspath))
        if item.get_closest_marker("gpu") is None or name not in _GPU_FILE_ORDER:
            return (0, module_pos[str(item.fspath)], 1 if item.name.split("[")[0] in _COLLECT_LAST_IN_MODULE else 0)
        return (1, _GPU_FILE_ORDER.index(name), 0)
    items.sort(key=key)          # stable: CPU tests keep their order and run first, GPU tests follow in the order above


@pytest.fixture(scope="session")
def state_dicts():
    """One deterministic state_dict per preset (oracle/weights.py, seed 5), built lazily."""
    import weights
    cache = {}

    def get(preset):
        if preset not in cache:
            cache[preset] = weights.make_state_dict(preset, seed=5)
        return cache[preset]
    return get
