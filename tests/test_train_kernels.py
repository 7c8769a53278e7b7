"""Op-level parity of the training kernels (egotap_b200/csrc/train_ops.cu, train_model.cu) with the op oracle, on
deliberately awkward shapes, through two execution vehicles:
  * "emu"  (CPU suite): the kernels' own SOURCE compiled by g++ against a CUDA-on-CPU shim (tests/cuda_emu: CTAs in
    sequence, threads as fibers with real __syncthreads / warp-shuffle semantics) -- checks indexing, layouts,
    reductions and arithmetic of the kernel code without a GPU
  * "cuda" (-m gpu): the same comparisons against the real kernels through the C ABI on the B200
    (tests/gpu_adapter.py mirrors the CPU test arguments on the device)
and, for "emu", end to end inside the training engine."""
import os
import sys

import pytest
import torch

import op_oracle
import train_oracle as tro
import weights
from egotap_b200 import training
from egotap_b200.synthetic import synthetic_heatmaps

sys.path.insert(0, os.path.dirname(__file__))

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "cuda_emu"))
import build_emu  # noqa: E402

BF16 = torch.bfloat16


# "emu-reversed" / "emu-shuffled": the same kernel source with the threads (and blocks) of every launch scheduled in
# descending / per-round reshuffled order.  CUDA promises no order between threads that no barrier separates, so the results
# must not change: a missing __syncthreads / __syncwarp or an in-place hazard between threads shows up in one of the orders.
SCHEDULES = {"emu": 0, "emu-reversed": 1, "emu-shuffled": 2}


@pytest.fixture(scope="module", params=list(SCHEDULES) + [pytest.param("cuda", marks=[pytest.mark.gpu, pytest.mark.timeout(600, method="thread")])])
def backends(request):
    if request.param in SCHEDULES:
        import ctypes
        pair = build_emu.make_backend()
        pair[0].name = "emu"
        pair[0].schedule = request.param
        lib = ctypes.CDLL(build_emu.build())
        lib.emu_set_schedule.argtypes = [ctypes.c_int, ctypes.c_ulonglong]
        lib.emu_set_schedule(SCHEDULES[request.param], 12345)
        yield pair
        lib.emu_set_schedule(0, 0)
        return
    from gpu_adapter import GpuOpAdapter
    yield GpuOpAdapter(), op_oracle.OracleBackend()


@pytest.fixture(scope="module")
def emu_backends():
    """the emulated-kernel backend on the default thread schedule (engine-level / host-logic tests run once, CPU only; their
    hardware counterparts are tests/test_zz_train_gpu.py)"""
    pair = build_emu.make_backend()
    pair[0].name = "emu"
    pair[0].schedule = "emu"
    return pair


def _pair(rows, cols, poison=True):
    mk = lambda: torch.full((rows, cols), float("nan"), dtype=BF16) if poison else torch.zeros((rows, cols), dtype=BF16)
    return mk(), mk()


def _same(a, b, tol=0.0):
    a, b = a.float(), b.float()
    assert torch.isnan(a).equal(torch.isnan(b)), "NaN (never-written) pattern differs"
    a, b = torch.nan_to_num(a), torch.nan_to_num(b)
    err = (a - b).abs().max().item()
    assert err <= tol * max(b.abs().max().item(), 1e-30), err


def _both(backends, name, make_args):
    """run op `name` on the emulated kernels and on the oracle with identical (cloned) arguments; returns both arg lists"""
    emu, orc = backends
    a1, a2 = make_args(), make_args()
    getattr(emu, name)(*a1)
    getattr(orc, name)(*a2)
    return a1, a2


@pytest.mark.parametrize("rows,cols,rows_in,rows_out", [(90, 128, 0, 0), (130, 192, 0, 0), (96, 64, 72, 48)])
def test_transpose_split(backends, rows, cols, rows_in, rows_out):
    torch.manual_seed(0)
    src_rows = rows if rows_out == 0 else (rows // rows_out) * rows_in
    ld = cols + 8
    src = torch.randn(src_rows, ld)
    pad = (rows + 63) // 64 * 64 + 64

    def args():
        rh, rl = _pair(rows, cols)
        th, tl = _pair(cols, pad + 2)
        return [src.clone(), rows, cols, ld, rows_in, rows_out, rh, rl, cols, th, tl, pad + 2, pad]
    a, b = _both(backends, "transpose_split", args)
    for i in (6, 7, 9, 10):
        _same(a[i], b[i])
    # row-major only / transposed only
    a, b = _both(backends, "transpose_split", lambda: args()[:6] + [None, None, 0] + args()[9:])
    _same(a[9], b[9])
    a, b = _both(backends, "transpose_split", lambda: args()[:9] + [None, None, 0, 0])
    _same(a[6], b[6])
    # fused forms: gelu'(u) factor and the column sums (bias gradient), with and without the bf16 outputs
    u = torch.randn(src_rows, ld) * 2
    scr = torch.zeros(((rows + 63) // 64) * cols)
    a, b = _both(backends, "transpose_split", lambda: args() + [u, torch.full((cols,), float("nan")), scr.clone()])
    for i in (6, 9):
        _same(a[i].float() + a[i + 1].float(), b[i].float() + b[i + 1].float(), 2e-5)
    _same(a[14], b[14], 2e-6)
    a, b = _both(backends, "transpose_split", lambda: args()[:6] + [None, None, 0, None, None, 0, 0, None,
                                                                    torch.full((cols,), float("nan")), scr.clone()])
    _same(a[14], b[14], 2e-6)


def test_transpose_bf16_grouped(backends):
    torch.manual_seed(1)
    frames, heads, tok, hd = 2, 3, 70, 64                 # rows not a multiple of 64: ragged tiles + padding
    src_h, src_l = torch.randn(frames * tok, heads * hd).to(BF16), torch.randn(frames * tok, heads * hd).to(BF16)
    pad = 128

    def args():
        dh, dl = _pair(frames * heads * hd, pad)
        return [src_h, src_l, tok, hd, heads * hd, heads, hd, frames, tok * heads * hd, dh, dl, pad, hd * pad, heads * hd * pad, pad]
    a, b = _both(backends, "transpose_bf16", args)
    _same(a[9], b[9]); _same(a[10], b[10])
    a, b = _both(backends, "transpose_bf16", lambda: [src_h, None] + args()[2:10] + [None] + args()[11:])
    _same(a[9], b[9])


@pytest.mark.parametrize("rows,cols,rows_in,rows_out", [(1000, 256, 0, 0), (5, 4096, 0, 0), (96, 128, 72, 48), (3, 36864, 0, 0)])
def test_colsum_and_reduce_partials(backends, rows, cols, rows_in, rows_out):
    torch.manual_seed(2)
    src_rows = rows if rows_out == 0 else (rows // rows_out) * rows_in
    src = torch.randn(src_rows, cols + 4)
    scr = torch.zeros(1 << 16)
    a, b = _both(backends, "colsum", lambda: [src, rows, cols, cols + 4, rows_in, rows_out, torch.zeros(cols), scr])
    _same(a[6], b[6], 2e-6)
    part = torch.randn(5, 1024)
    a, b = _both(backends, "reduce_partials", lambda: [part, 5, 1024, torch.zeros(1024)])
    _same(a[3], b[3], 1e-6)
    # tall and narrow (the column sums behind a bias gradient: thousands of per-tile rows): the 16-row-lane kernel, ragged n
    tall = torch.randn(203, 1000)
    a, b = _both(backends, "reduce_partials", lambda: [tall, 203, 1000, torch.full((1000,), float("nan"))])
    _same(a[3], b[3], 5e-6)


def test_gelu(backends):
    u = torch.linspace(-9, 9, 4096)
    a, b = _both(backends, "gelu_fwd", lambda: [u, 4096, *_pair(1, 4096)])
    _same(a[2].float() + a[3].float(), b[2].float() + b[3].float(), 2e-5)     # hi+lo carries ~16 mantissa bits
    dg = torch.randn(4096)
    a, b = _both(backends, "gelu_bwd", lambda: [dg.clone(), u, 4096])
    _same(a[0], b[0], 2e-6)


@pytest.mark.parametrize("frames,rows_in,rows_out,acc", [(3, 8, 8, 1), (2, 12, 7, 0), (40, 10, 10, 1)])
def test_layernorm_bwd(backends, frames, rows_in, rows_out, acc):
    torch.manual_seed(3)
    x = torch.randn(frames * rows_in, 1024) * 2 + 0.3
    dy = torch.randn(frames * rows_out, 1024)
    w = torch.rand(1024) + 0.5
    dx0 = torch.randn(frames * rows_in, 1024)
    scr = torch.zeros(1 << 20)
    a, b = _both(backends, "layernorm_bwd", lambda: [dy, x, w, frames, rows_in, rows_out, 1e-12, dx0.clone(), acc,
                                                      torch.zeros(1024), torch.zeros(1024), scr])
    _same(a[7], b[7], 1e-5)
    _same(a[9], b[9], 1e-5)
    _same(a[10], b[10], 1e-5)


def test_softmax_bwd(backends):
    torch.manual_seed(4)
    rows = 21
    S, dP = torch.randn(rows, 576) * 3, torch.randn(rows, 576)
    a, b = _both(backends, "softmax_bwd", lambda: [S, dP, rows, 576, 0.0883883, *_pair(rows, 576), *_pair(rows, 576)])
    for i in (5, 7):
        _same(a[i].float() + a[i + 1].float(), b[i].float() + b[i + 1].float(), 1e-5)


@pytest.mark.parametrize("rows,cols,J", [(90, 2048, 0), (60, 128, 15), (68, 128, 17), (7, 512, 0)])
def test_batchnorm_train(backends, rows, cols, J):
    torch.manual_seed(5)
    y = torch.randn(rows, cols) * 1.7 + 0.4
    gamma, beta = torch.rand(cols) + 0.5, torch.randn(cols) * 0.1
    scr = torch.zeros(1 << 18)

    rm0, rv0 = torch.randn(cols), torch.rand(cols) + 0.5

    def stats_args():
        return [y, rows, cols, gamma, beta, rm0.clone(), rv0.clone(), torch.tensor(3), 0.1, 1e-5,
                torch.zeros(cols), torch.zeros(cols), torch.zeros(cols), torch.zeros(cols), scr]
    a, b = _both(backends, "bn_stats", stats_args)
    for i in (5, 6, 10, 11, 12, 13):
        _same(a[i], b[i], 2e-6)
    assert int(a[7]) == int(b[7]) == 4
    mean, rstd, scale, shift = b[10], b[11], b[12], b[13]
    out_rows, out_cols = (rows // 2 if J else rows), (2 * cols + 256 if J else cols + 32)

    def apply_args():
        return [y, rows, cols, scale, shift, *_pair(out_rows, out_cols), out_cols, torch.full((out_rows, out_cols), float("nan")),
                out_cols, J, 32 if not J else 256]
    a, b = _both(backends, "bn_apply", apply_args)
    # y*scale + shift contracts to one FMA on the device and is two roundings in the oracle: the fp32 value may differ by an ulp,
    # which can flip the bf16 rounding of `hi` alone -- the pair is compared as hi + lo (~16 mantissa bits), like every other
    # pair that is the result of arithmetic
    _same(a[5].float() + a[6].float(), b[5].float() + b[6].float(), 2e-5); _same(a[8], b[8], 1e-6)
    da = torch.randn(rows, cols)
    a, b = _both(backends, "bn_bwd", lambda: [da.clone(), y, rows, cols, scale, shift, mean, rstd, torch.zeros(cols),
                                               torch.zeros(cols), scr])
    _same(a[0], b[0], 2e-5); _same(a[8], b[8], 2e-5); _same(a[9], b[9], 2e-5)
    if J:
        dE = torch.randn(out_rows, 512)
        a, b = _both(backends, "regroup_gather", lambda: [dE, 512, 256, rows // (2 * J), J, cols, torch.zeros(rows, cols)])
        _same(a[6], b[6])


@pytest.mark.parametrize("J,B", [(15, 5), (17, 3)])
def test_pu_cell_forward_and_backward(backends, J, B):
    torch.manual_seed(6)
    H = 512
    g_ld, f_ld = 5 * H, 5 * H                              # the layer-1 layout: [F | G] interleaved per row
    FG = torch.randn(B * J, 5 * H)
    dOut = torch.randn(B * J, H)

    def fwd(t, st):
        return [FG[:, H:], J * g_ld, g_ld, FG, J * f_ld, f_ld, st["C"], st["H"], st["hh"], st["hl"], st["gh"], st["gl"], t, J, B]

    def state():
        hh, hl = _pair(B * J, H); gh, gl = _pair(B * J, H, poison=False)
        return dict(C=torch.full((B * J, H), float("nan")), H=torch.full((B * J, H), float("nan")), hh=hh, hl=hl, gh=gh, gl=gl)
    emu, orc = backends
    s1, s2 = state(), state()
    for t in range(J):
        emu.pu_cell_fwd(*fwd(t, s1)); orc.pu_cell_fwd(*fwd(t, s2))
    _same(s1["C"], s2["C"], 2e-6); _same(s1["H"], s2["H"], 2e-6)
    for hi, lo in (("hh", "hl"), ("gh", "gl")):               # bf16 pairs are compared as hi + lo (~16 mantissa bits)
        _same(s1[hi].float() + s1[lo].float(), s2[hi].float() + s2[lo].float(), 2e-5)
    st = s2

    def bstate():
        return dict(dhg=torch.randn(B, H), dc=torch.randn(B, H), dFG=torch.full((B * J, 5 * H), float("nan")), dg=_pair(B, 4 * H))
    torch.manual_seed(7); b1 = bstate()
    torch.manual_seed(7); b2 = bstate()
    for t in range(J - 1, -1, -1):
        for be, bs in ((emu, b1), (orc, b2)):
            be.pu_cell_bwd(FG[:, H:], J * g_ld, g_ld, FG, J * f_ld, f_ld, st["C"], st["H"], dOut, bs["dhg"], bs["dc"],
                           bs["dFG"][:, H:], J * g_ld, g_ld, bs["dFG"], J * f_ld, f_ld, bs["dg"][0], bs["dg"][1], t, J, B)
        _same(b1["dc"], b2["dc"], 1e-5)
        _same(b1["dg"][0].float() + b1["dg"][1].float(), b2["dg"][0].float() + b2["dg"][1].float(), 1e-4)
        nxt = torch.randn(B, H)                             # stands in for the dgates . W_hh GEMM of the real backward
        b1["dhg"], b2["dhg"] = nxt.clone(), nxt.clone()
    _same(b1["dFG"], b2["dFG"], 1e-5)


def test_bridge_gate_bwd(backends):
    torch.manual_seed(8)
    rows = 45
    F0, E, dE, = torch.randn(rows, 768), torch.randn(rows, 512), torch.randn(rows, 512)
    a, b = _both(backends, "pu_bridge_gate_bwd", lambda: [dE.clone(), 512, F0, 768, 512, E, 256, rows, torch.full((rows, 768), float("nan")), 768])
    _same(a[0], b[0], 5e-6); _same(a[8], b[8], 5e-6)     # a few ulp: expf / division / FMA contraction differ between libm and the device


@pytest.mark.parametrize("preset,J,B", [("UnrealEgo", 15, 7), ("EgoCap", 17, 4)])
def test_head_bwd_embed_grads_loss_adamw(backends, preset, J, B):
    torch.manual_seed(9)
    gh = preset == "UnrealEgo"
    nj = J + 1 if gh else J
    e, skel = torch.randn(B * J, 512), torch.randn(B * J, 512)
    Wp, Wg = torch.randn(3, 768), (torch.randn(6, J * 512) if gh else None)
    dpose = torch.randn(B, nj, 3)
    scr = torch.zeros(1 << 20)

    def args():
        return [dpose, e, 512, skel, Wp, Wg, B, J, torch.full((B * J, 512), float("nan")), 512, torch.full((B * J, 512), float("nan")),
                torch.zeros(3, 768), torch.zeros(3), torch.zeros(6, J * 512) if gh else None, torch.zeros(6) if gh else None, scr]
    a, b = _both(backends, "head_bwd", args)
    for i in (8, 10, 11, 12) + ((13, 14) if gh else ()):
        _same(a[i], b[i], 1e-5)
    dpp = torch.randn(576, 1024)
    a, b = _both(backends, "embed_grads", lambda: [dpp, 6, 2 * J, torch.zeros(576, 1024), torch.zeros(1024)])
    _same(a[3], b[3]); _same(a[4], b[4], 1e-6)
    pred, gt = torch.randn(B, nj, 3) * 10, torch.randn(B, nj, 3) * 10
    a, b = _both(backends, "pose_loss", lambda: [pred, gt, B, nj, training.KINEMATIC_PARENTS[preset], not gh, 0.1, -0.01,
                                                  torch.zeros(4), torch.zeros(B, nj, 3), scr])
    _same(a[8][:3], b[8][:3], 1e-5); _same(a[9], b[9], 1e-5)
    ps = [torch.randn(n) for n in (3, 1000, 70001)] + [torch.randn(1001)[1:]]      # last: 4-byte aligned only -> scalar path
    gs, ms, vs = [torch.randn_like(p) for p in ps], [torch.rand_like(p) * 0.1 for p in ps], [torch.rand_like(p) * 0.01 for p in ps]
    a, b = _both(backends, "adamw", lambda: [[p.clone() for p in ps], gs, [m.clone() for m in ms], [v.clone() for v in vs], 3, 1e-3,
                                              0.9, 0.999, 1e-4, 0.01])
    for i in (0, 2, 3):
        for x, y in zip(a[i], b[i]):
            _same(x, y, 1e-5)


def test_engine_on_emulated_kernels_matches_autograd(emu_backends):
    """the whole training step with every training op running on the emulated kernel source (GEMM / attention /
    LayerNorm / ingest / head from the oracle): gradients vs torch.autograd on the restated forward"""
    emu, _ = emu_backends
    preset, batch = "UnrealEgo", 1
    sd = weights.make_state_dict(preset, seed=5)
    params = {k: v.clone().contiguous() for k, v in sd.items()}
    eng = training.TrainEngine(preset, params, precision="bf16x3", backend=emu)
    x = synthetic_heatmaps(preset, batch, seed=17, kind="gauss")
    gt = torch.randn(batch, 16, 3, generator=torch.Generator().manual_seed(19)) * 20
    ref_loss, ref_sd, _, ref_grads = tro.train_step(sd, x, gt, preset)
    eng.forward(x.clone())
    loss = eng.loss_and_grad(gt.clone())
    assert abs(float(loss[0]) - float(ref_loss)) < 5e-5 * max(1.0, abs(float(ref_loss)))
    grads = eng.backward()
    assert not torch.isnan(eng.flat_grad).any()
    for k, g_ref in ref_grads.items():
        if g_ref is None or g_ref.abs().max() < 1e-7:
            continue
        a, b = grads[k].flatten().double(), g_ref.flatten().double()
        cos = float((a @ b) / (a.norm() * b.norm()))
        assert cos > 0.999, (k, cos)
        assert abs(float(a.norm() / b.norm()) - 1) < 3e-2, k
    eng.adamw_step()
    # AdamW turns every gradient into a step of ~ +-lr, so the update is compared only downstream of the LeakyReLUs
    # (a single flipped LeakyReLU' factor at z ~ 0 legitimately moves ViT-side steps by O(lr))
    for k in ("pose_mlp.pose_fcs.0.weight", "skel_sequential_layer.lstm_custom.layers.1.h2h.weight"):
        upd_ref, upd = ref_sd[k] - sd[k], params[k] - sd[k]
        assert (upd - upd_ref).abs().max().item() <= 5e-2 * upd_ref.abs().max().item() + 2e-7, k


def test_recorded_step_replays_identically():
    """TrainEngine records the library calls of one step and replays them afterwards (no Python orchestration on the
    hot path): two steps (record, replay) must leave exactly the same weights as two steps without.  Host-side logic; the CUDA
    engine runs the same code path in tests/test_zz_train_gpu.py"""
    emu, _ = build_emu.make_backend(all_oracle=True)      # record / replay is host logic: ops served by the oracle
    preset, batch = "EgoCap", 1
    sd = weights.make_state_dict(preset, seed=5)
    x = synthetic_heatmaps(preset, batch, seed=17, kind="gauss")
    gt = torch.randn(batch, 17, 3, generator=torch.Generator().manual_seed(19)) * 20
    results = []
    for use_tape in (True, False):
        params = {k: v.clone().contiguous() for k, v in sd.items()}
        eng = training.TrainEngine(preset, params, precision="bf16", backend=emu)
        assert eng.use_tape
        eng.use_tape = use_tape
        losses = [float(eng.train_step(x.clone(), gt.clone())[0]) for _ in range(2)]
        if use_tape:
            assert eng._tape is not None and len(eng._tape[0]) > 300
        results.append((losses, {k: params[k].clone() for k in ("pose_mlp.pose_fcs.0.weight",
                                                                "pos_heatmap_encoder.vit.encoder.layer.0.output.dense.weight",
                                                                "rot_heatmap_encoder.fc1.bn.running_var")}))
    assert results[0][0] == results[1][0] and results[0][0][0] != results[0][0][1]
    for k in results[0][1]:
        assert torch.equal(results[0][1][k], results[1][1][k]), k


def test_emulation_traps_misaligned_vector_access():
    """the emulation library is built with -fsanitize=alignment: a float4 access through a pointer that is only 4-byte
    aligned must abort (run in a subprocess) -- so the green emulated runs above also certify the alignment of every
    vector load / store the engine's call sites produce"""
    import subprocess
    code = (
        "import ctypes as C, sys; sys.path.insert(0, %r); import build_emu\n"
        "L = C.CDLL(build_emu.build())\n"
        "buf = (C.c_float * 64)(); out = (C.c_uint16 * 64)()\n"
        "base = C.addressof(buf)\n"
        "L.egotap_b200_gelu_fwd.argtypes = [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]\n"
        "assert L.egotap_b200_gelu_fwd(base, 16, C.addressof(out), None, None) == 0\n"
        "print('aligned ok', flush=True)\n"
        "L.egotap_b200_gelu_fwd(base + 4, 16, C.addressof(out), None, None)\n"
        "print('misaligned survived', flush=True)\n" % os.path.join(os.path.dirname(__file__), "cuda_emu"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "aligned ok" in r.stdout and "misaligned survived" not in r.stdout and r.returncode != 0, (r.stdout, r.stderr[-400:])
    assert "misaligned address" in r.stderr


@pytest.mark.parametrize("preset,J", [("UnrealEgo", 15), ("EgoCap", 17)])
def test_round1_bandwidth_kernels(backends, preset, J):
    """the round-1 kernels of csrc/kernels.cu that the training engine reuses (ingest, LayerNorm with row compaction,
    regression head, packing helpers): same vehicles, same oracle"""
    torch.manual_seed(11)
    B = 2
    x = torch.rand(B, 6 * J, 64, 64)
    a, b = _both(backends, "ingest", lambda: [x, J, *_pair(B * 2 * J * 16, 256), *_pair(B * 2 * J, 8192)])
    for i in (2, 3, 4, 5):
        _same(a[i], b[i])
    h = torch.randn(B * 576, 1024) * 2 + 0.5
    w, bias = torch.rand(1024) + 0.5, torch.randn(1024) * 0.1
    live = 2 * J * 16
    a, b = _both(backends, "layernorm", lambda: [h, w, bias, B, 576, live, 1e-12, *_pair(B * live, 1024),
                                                 torch.full((B * live, 1024), float("nan"))])
    _same(a[9], b[9], 2e-6)
    _same(a[7].float() + a[8].float(), b[7].float() + b[8].float(), 2e-5)
    e, skel = torch.randn(B * J, 512), torch.randn(B * J, 512)
    Wp, bp = torch.randn(3, 768), torch.randn(3)
    gh = preset == "UnrealEgo"
    Wg, bg = (torch.randn(6, J * 512), torch.randn(6)) if gh else (None, None)
    a, b = _both(backends, "head", lambda: [e, 512, skel, Wp, bp, Wg, bg, B, J, torch.full((B, J + 1 if gh else J, 3), float("nan"))])
    _same(a[9], b[9], 1e-5)
    pos, mask = torch.randn(1, 576, 1024), torch.randn(1, 1, 1024)
    a, b = _both(backends, "pos_permute", lambda: [pos, mask, 6, 2 * J, torch.full((576, 1024), float("nan")),
                                                   torch.full((576 - live, 1024), float("nan"))])
    _same(a[4], b[4]); _same(a[5], b[5], 1e-6)
    hid = torch.randn(B * 576, 1024)
    a, b = _both(backends, "fill_dummy", lambda: [hid.clone(), b[5], B, 576, live])
    _same(a[0], b[0])
    F0, E = torch.randn(B * J, 768), torch.randn(B * J, 512)
    a, b = _both(backends, "pu_bridge_gate", lambda: [F0, 768, 512, E, 512, 256, B * J, *_pair(B * J, 512)])
    _same(a[7].float() + a[8].float(), b[7].float() + b[8].float(), 2e-5)
    src = torch.randn(40, 72)
    a, b = _both(backends, "split2d", lambda: [src, 40, 64, 72, *_pair(40, 128), 128])
    _same(a[4], b[4]); _same(a[5], b[5])
    v1, v2, v3 = torch.randn(100), torch.randn(100), torch.randn(100)
    a, b = _both(backends, "add3", lambda: [v1, v2, v3, torch.zeros(100), 100])
    _same(a[3], b[3], 1e-6)


def test_gpu_adapter_mirroring_logic_on_cpu():
    """tests/gpu_adapter.py (the vehicle of every [cuda] op test) exercised without a GPU: emulation backend underneath,
    "device" = a cloned CPU buffer.  Aliasing views of one storage must stay aliased on the mirror and every storage must
    be copied back."""
    from gpu_adapter import GpuOpAdapter
    emu, orc = build_emu.make_backend()
    ad = GpuOpAdapter(backend=emu, to_device=lambda t: t.clone(), synchronize=lambda: None)
    torch.manual_seed(3)
    J, B, H = 15, 3, 512
    FG = torch.randn(B * J, 5 * H)
    dOut = torch.randn(B * J, H)
    C_, Hh = torch.randn(B * J, H), torch.randn(B * J, H)
    res = []
    for be_ in (ad, orc):
        dFG = torch.full((B * J, 5 * H), float("nan"))
        dhg, dc = torch.randn(B, H, generator=torch.Generator().manual_seed(1)), torch.randn(B, H, generator=torch.Generator().manual_seed(2))
        dg = _pair(B, 4 * H)
        # dG and dF are two views of ONE buffer (the layer-1 layout), as the engine passes them
        be_.pu_cell_bwd(FG[:, H:], J * 5 * H, 5 * H, FG, J * 5 * H, 5 * H, C_, Hh, dOut, dhg, dc, dFG[:, H:], J * 5 * H, 5 * H, dFG,
                        J * 5 * H, 5 * H, dg[0], dg[1], J - 2, J, B)
        res.append((dFG, dc, dg[0].float() + dg[1].float()))
    for a, b in zip(res[0], res[1]):
        _same(a, b, 1e-4)
    ps = [torch.randn(70), torch.randn(1001)[1:]]
    gs, ms, vs = [torch.randn_like(p) for p in ps], [torch.zeros_like(p) for p in ps], [torch.zeros_like(p) for p in ps]
    before = [p.clone() for p in ps]
    ad.adamw(ps, gs, ms, vs, 1, 1e-3, 0.9, 0.999, 1e-4, 0.0)
    assert all(not torch.equal(p, b) for p, b in zip(ps, before)) and all(float(m.abs().max()) > 0 for m in ms)
