"""Op-level parity (through the C ABI) of the non-GEMM kernels against the oracle's corresponding functions."""
import pytest
import torch
import torch.nn.functional as F

import egotap_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
def test_ingest_layouts(preset):
    """reference model/net_architecture.py:688-694 (split / re-layout) and :375-383 + modeling_vit.py:195 (patchify)."""
    from egotap_b200 import capi, synthetic_heatmaps
    x = synthetic_heatmaps(preset, 3, seed=5, kind="uniform")
    ph, pl, lh, ll = capi.ingest(x.cuda(), preset)
    pos, rot = orc.split_input(x, preset)
    B, n = pos.shape[0], pos.shape[1]
    # patches of heatmap n in (pr, pc) order, each flattened (py, px) like the conv kernel
    want_p = pos.reshape(B, n, 4, 16, 4, 16).permute(0, 1, 2, 4, 3, 5).reshape(B * n * 16, 256)
    want_l = rot.reshape(B * n, 2 * 64 * 64)
    got_p = ph.float().cpu() + pl.float().cpu()
    got_l = lh.float().cpu() + ll.float().cpu()
    assert (got_p - want_p).abs().max() < 1e-4 * want_p.abs().max()      # hi+lo carries ~16 mantissa bits
    assert (got_l - want_l).abs().max() < 1e-4 * max(want_l.abs().max(), 1e-6)
    assert torch.equal(ph.cpu(), want_p.to(torch.bfloat16))              # hi is exactly the bf16 rounding


@pytest.mark.parametrize("rows_out", [576, 480, 544])
def test_layernorm_and_compaction(rows_out):
    from egotap_b200 import capi
    torch.manual_seed(0)
    frames = 3
    x = torch.randn(frames * 576, 1024) * 3 + 0.7
    w, b = torch.rand(1024) + 0.5, torch.randn(1024) * 0.1
    hi, lo, f32 = capi.layernorm(x.cuda(), w.cuda(), b.cuda(), frames, 576, rows_out, 1e-12)
    ref = F.layer_norm(x.double(), (1024,), w.double(), b.double(), 1e-12).view(frames, 576, 1024)[:, :rows_out].reshape(-1, 1024)
    assert (f32.cpu().double() - ref).abs().max() < 2e-5
    assert ((hi.float() + lo.float()).cpu().double() - ref).abs().max() < 2e-4


@pytest.mark.parametrize("frames,J", [(5, 15), (300, 17)])
@pytest.mark.parametrize("x3", [True, False])
def test_propagation_layer_against_cell_recurrence(frames, J, x3):
    """pu_chain_kernel vs the oracle's cell recurrence (reference custom_cells.py:94-120) for one layer, given the
    batched x-side terms.  300 frames = two batch groups of the persistent kernel, ragged second group."""
    from egotap_b200 import capi
    torch.manual_seed(1)
    H = 512
    W = torch.randn(4 * H, H) / H ** 0.5
    G = torch.randn(frames * J, 4 * H)
    Fg = torch.randn(frames * J, H)
    out = capi.pu_chain(W.cuda(), G.cuda(), Fg.cuda(), frames, J, capi.PREC_BF16X3 if x3 else capi.PREC_BF16)
    Wd = (W if x3 else W.to(torch.bfloat16).float()).double()
    h = torch.zeros(frames, H, dtype=torch.float64); c = torch.zeros_like(h)
    Gd, Fd = G.double().view(frames, J, -1), Fg.double().view(frames, J, -1)
    ref = []
    for t in range(J):
        hg = torch.sigmoid(Fd[:, t]) * h
        if not x3:
            hg = hg.float().to(torch.bfloat16).double()
        g = Gd[:, t] + hg @ Wd.t()
        fg, ig, cg, og = g.chunk(4, 1)
        c = c * torch.sigmoid(fg) + torch.sigmoid(ig) * torch.tanh(cg)
        h = torch.sigmoid(og) * torch.tanh(c)
        ref.append(h)
    ref = torch.stack(ref, 1).reshape(frames * J, H)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < (2e-5 if x3 else 2e-3), err


@pytest.mark.parametrize("preset", ["UnrealEgo", "EgoCap"])
def test_head_kernel(preset, state_dicts):
    """reference model/net_architecture.py:732-751: per-joint Linear(768->3), global offset, head joint last."""
    from egotap_b200 import capi
    sd = state_dicts(preset)
    g = orc.geometry(preset)
    J, B = g["J"], 4
    torch.manual_seed(2)
    pe, re_, skel = torch.randn(B, J, 256), torch.randn(B, J, 256), torch.randn(B, J, 512)
    e = torch.cat([pe, re_], -1).reshape(B * J, 512).contiguous()
    Wg = sd.get("global_mlp.pose_fcs.0.weight"); bg = sd.get("global_mlp.pose_fcs.0.bias")
    pose = capi.head(e.cuda(), skel.reshape(B * J, 512).cuda(), sd["pose_mlp.pose_fcs.0.weight"].cuda(),
                     sd["pose_mlp.pose_fcs.0.bias"].cuda(), None if Wg is None else Wg.cuda(),
                     None if bg is None else bg.cuda(), B, J)
    per_joint = torch.cat([pe, skel], -1).reshape(B * J, 768)
    ref = F.linear(per_joint, sd["pose_mlp.pose_fcs.0.weight"], sd["pose_mlp.pose_fcs.0.bias"]).view(B, J, 3)
    if Wg is not None:
        o = F.linear(skel.reshape(B, J * 512), Wg, bg)
        ref = torch.cat([ref + o[:, None, :3], o[:, None, 3:]], 1)
    assert pose.shape == ref.shape
    assert (pose.cpu() - ref).abs().max() < 2e-5
