"""Evaluation metrics: oracle vs reference (golden + live, CPU) and the fused CUDA kernel vs
the oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

import metrics_oracle as mo
import ref_shim

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_metrics.npz")


def _cases():
    g = torch.Generator().manual_seed(11)
    gt = torch.randn(300, 17, 3, generator=g) * 25
    pred = gt + torch.randn(300, 17, 3, generator=g) * 4
    pred[7] = gt[7] * torch.tensor([1.0, 1.0, -1.0])             # mirrored
    pred[8] = gt[8] * 0.5 + 3.0                                    # pure similarity: PA error must vanish
    return pred, gt


def test_oracle_matches_reference_golden():
    d = np.load(GOLD)
    m, pa = mo.pose_metrics(torch.from_numpy(d["pred"]), torch.from_numpy(d["gt"]))
    np.testing.assert_allclose(m.numpy(), d["mpjpe_mm"], rtol=2e-6, atol=1e-4)
    np.testing.assert_allclose(pa.numpy(), d["pa_mpjpe_mm"], rtol=2e-6, atol=1e-4)


@pytest.mark.skipif(ref_shim.reference_root() is None, reason="reference tree not present")
def test_oracle_matches_live_reference():
    ref_shim.import_reference()
    from utils.loss import LossFuncMPJPE
    from utils.util import batch_compute_similarity_transform_torch
    pred, gt = _cases()
    S = batch_compute_similarity_transform_torch(pred, gt)
    lf = LossFuncMPJPE()
    ref_pa = torch.stack([lf(S[i], gt[i]) * 10 for i in range(pred.shape[0])])
    m, pa = mo.pose_metrics(pred, gt)
    assert (pa.float() - ref_pa).abs().max() < 2e-3
    assert pa[8] < 1e-4        # exact similarity up to fp32 rounding of the inputs


@pytest.mark.gpu
@pytest.mark.parametrize("joints", [16, 17])
def test_kernel_matches_oracle(joints):
    from egotap_b200.evaluate import compute_metrics, pose_metrics
    pred, gt = _cases()
    pred, gt = pred[:, :joints].contiguous(), gt[:, :joints].contiguous()
    m, pa = pose_metrics(pred.cuda(), gt.cuda())
    rm, rpa = mo.pose_metrics(pred, gt)
    assert (m.cpu().double() - rm).abs().max() < 1e-4
    assert (pa.cpu().double() - rpa).abs().max() < 1e-4

    class Avg:                                   # the reference's RunningAverageDict protocol: .update(dict)
        def __init__(self): self.n, self.s = 0, {}
        def update(self, d):
            self.n += 1
            for k, v in d.items(): self.s[k] = self.s.get(k, 0.0) + float(v)
    avg = Avg()
    mp, pap = compute_metrics(pred.cuda(), gt.cuda(), avg)
    assert avg.n == pred.shape[0] and abs(avg.s["pa_mpjpe"] / avg.n - rpa.mean().item()) < 1e-4
    assert mp.device.type == "cpu" and pap.shape == (pred.shape[0],)


@pytest.mark.gpu
def test_kernel_matches_reference_golden():
    from egotap_b200.evaluate import pose_metrics
    d = np.load(GOLD)
    m, pa = pose_metrics(torch.from_numpy(d["pred"]).cuda(), torch.from_numpy(d["gt"]).cuda())
    np.testing.assert_allclose(m.cpu().numpy(), d["mpjpe_mm"], rtol=1e-5, atol=2e-4)
    np.testing.assert_allclose(pa.cpu().numpy(), d["pa_mpjpe_mm"], rtol=1e-5, atol=2e-4)


def test_saved_result_files_are_byte_compatible(tmp_path):
    """pred_pose.npy / gt_<dataset>_pose.npy / input_<dataset>_paths.npy / input_paths.pkl exactly as the reference's
    test_evaluate(save_result=True) writes them (utils/evaluate.py:127-144)"""
    import io
    import os
    import pickle

    import numpy as np
    from egotap_b200.evaluate import save_predictions
    rng = np.random.default_rng(0)
    preds = [rng.normal(size=(4, 16, 3)).astype(np.float32), rng.normal(size=(3, 16, 3)).astype(np.float32)]
    gts = [rng.normal(size=(4, 16, 3)).astype(np.float32), rng.normal(size=(3, 16, 3)).astype(np.float32)]
    paths = [["/d/a/%d.npy" % i for i in range(4)], ["/d/b/%d.npy" % i for i in range(3)]]
    save = tmp_path / "exp" / "results"
    save.mkdir(parents=True)
    save_predictions([torch.from_numpy(p) for p in preds], gts, paths, str(save), "/data/UnrealEgoData/")
    # the reference's statements, restated
    want_pred, want_gt = np.concatenate(preds, 0), np.concatenate(gts, 0)
    want_paths = np.concatenate(paths, 0).reshape(-1, 1)

    def npy_bytes(a):
        b = io.BytesIO()
        np.save(b, a)
        return b.getvalue()
    assert (save / "pred_pose.npy").read_bytes() == npy_bytes(want_pred)
    assert (save.parent / "gt_unrealegodata_pose.npy").read_bytes() == npy_bytes(want_gt)
    assert (save.parent / "input_unrealegodata_paths.npy").read_bytes() == npy_bytes(want_paths)
    assert pickle.load(open(save / "input_paths.pkl", "rb")).tolist() == want_paths.tolist()
    assert np.load(save / "pred_pose.npy").shape == (7, 16, 3)
