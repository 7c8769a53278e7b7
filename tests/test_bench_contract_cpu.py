"""bench.py's reference arm (the part of the benchmark contract that runs without a GPU): one JSON line with the agreed
keys, for the default lifting workload and for the training workload; the torchrun contract (ranks != 0 print nothing)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _run(extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"] + extra,
                       capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


@pytest.mark.parametrize("extra,metric", [([], "stereo_frames_per_sec_heatmap_to_3d"),
                                          (["--workload", "train"], "train_frames_per_sec_lifting_net")])
def test_reference_arm_prints_one_contract_line(extra, metric):
    lines = _run(extra)
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == "frames/s" and d["value"] > 0
    # the lifting arm times the unmodified reference module when its tree is reachable (this container), else the oracle port
    has_ref = os.path.isfile("/root/reference/model/net_architecture.py") or os.environ.get("EGOTAP_REF")
    want_kind = "reference" if (has_ref and not extra) else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["higher_is_better"] is True


def test_reference_arm_other_ranks_stay_silent():
    assert _run([], env=dict(RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")) == []
