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


# GPU files in the order they are run: the inference path (hardware-verified in round 1) first, then the kernels that
# were brought up on the CPU emulation after the round's GPU budget was spent (first hardware run = the round-end tier).
# With -x a failure among the latter must not hide the parity tests of the path BASELINE.json names.
_GPU_FILE_ORDER = ["test_gemm_gpu.py", "test_ops_gpu.py", "test_lifting_gpu.py", "test_metrics.py", "test_heatmap_net.py",
                   "test_gt_heatmaps.py", "test_train_kernels.py", "test_zz_train_gpu.py", "test_zzz_graph_inference_gpu.py",
                   "test_zzz_attention_wide_gpu.py", "test_zzz_epilogue_coalesced_gpu.py"]


# CPU tests that run in a background child process started with their module (tests/test_tensorcore_emu.py): collected last
# within the module so that the wait for the child overlaps with the module's in-process tests
_COLLECT_LAST_IN_MODULE = {"test_whole_training_step_on_product_kernel_source", "test_engine_with_persistent_bptt_matches_per_joint_path",
                           "test_whole_inference_path_on_product_source"}


# GPU test files whose kernels have NOT run on hardware yet (written after round 1's GPU budget was spent; verified on the CPU
# emulation only).  Their GPU cases are quarantined as NON-strict expected failures: they still run, their outcome is reported
# (XPASS / XFAIL) and recorded under gpurun_out/, but a first-run failure neither stops a `-x` run nor paints the parity suite
# of the path BASELINE.json names red.  Remove a file from this list once its first hardware run is green (round 2, job scripts
# tools/gpu_job_r2a.sh / r2b.sh); EGOTAP_STRICT_UNVERIFIED=1 runs them as ordinary tests.
_UNVERIFIED_ON_HARDWARE = ["test_gt_heatmaps.py", "test_train_kernels.py", "test_zz_train_gpu.py", "test_zzz_attention_wide_gpu.py",
                           "test_zzz_epilogue_coalesced_gpu.py", "test_zzz_graph_inference_gpu.py"]


def pytest_collection_modifyitems(config, items):
    if os.environ.get("EGOTAP_STRICT_UNVERIFIED") != "1":
        for item in items:
            if item.get_closest_marker("gpu") is not None and os.path.basename(str(item.fspath)) in _UNVERIFIED_ON_HARDWARE:
                item.add_marker(pytest.mark.xfail(strict=False, reason="first hardware run pending (kernel verified on the CPU "
                                                  "emulation only; see tests/conftest.py _UNVERIFIED_ON_HARDWARE)"))
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
