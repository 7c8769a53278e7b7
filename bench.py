"""Benchmark of the EgoTAP heatmap->3D lifting path (BASELINE.json metric: stereo frames/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl egotap_b200|reference]
                    [--preset UnrealEgo|EgoCap] [--batch B] [--precision bf16x3|bf16]

A step = one forward pass of the lifting network over one batch of synthetic UnrealEgo-shaped
heatmaps (default: BASELINE.json configs[1], UnrealEgo preset, batch 256 per GPU, fp32-parity mode).
N > 1: one process per GPU under torchrun, frames sharded (weak scaling: 256 frames per GPU), the
only collective is the final gather of the poses.  Rank 0 prints ONE JSON line.

`--impl reference` times the reference's algorithm on the host cores (the CPU oracle port in
oracle/, since the reference is Python and does not travel to the GPU box), bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

FLOP_PER_FRAME = {"UnrealEgo": 51.161e9, "EgoCap": 51.600e9}     # BASELINE.md section 2 (algorithmic, 2*MAC)
METRIC = "stereo_frames_per_sec_heatmap_to_3d"
CPU_SAMPLE_BATCH = 16                                             # BASELINE.json configs[0]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm_gbs=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["unavailable"])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        out = dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx) if mx else None, reasons=reasons, samples=len(sm))

        def num(r, i):
            try:
                return float(r[i])
            except (IndexError, ValueError):
                return None
        # board power as nvidia-smi reports it (an average over about the last second, i.e. longer than a short timed region:
        # a lower bound of the power inside it; tools/power_probe.py measures sustained loops -- DESIGN section 9.2)
        pw = sorted(v for v in (num(r, 6) for r in self.rows) if v is not None)
        lim = [v for v in (num(r, 7) for r in self.rows) if v is not None]
        if pw:
            out.update(power_w_smi_avg=round(pw[len(pw) // 2], 1), power_limit_w=round(max(lim), 1) if lim else None)
        return out


def time_oracle_cpu(preset, sd, steps, warmup, batch=CPU_SAMPLE_BATCH):
    """The reference's algorithm (CPU oracle port, fp32, all host threads) on a bounded sample."""
    import torch
    import egotap_oracle as orc
    from egotap_b200.synthetic import synthetic_heatmaps
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core this process may run on
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    x = synthetic_heatmaps(preset, batch, seed=1234, kind="gauss")
    with torch.no_grad():
        for _ in range(warmup):
            orc.forward(sd, x, preset)
        t0 = time.perf_counter()
        for _ in range(steps):
            pose = orc.forward(sd, x, preset)
        dt = (time.perf_counter() - t0) / steps
    return dict(fps=batch / dt, ms_per_step=dt * 1e3, cores=torch.get_num_threads(), batch=batch, pose=pose, x=x)


def time_reference_module_cpu(preset, sd, steps, warmup, batch=CPU_SAMPLE_BATCH):
    """The UNMODIFIED reference module (model/net_architecture.py:682, imported through oracle/ref_shim.py) on the host cores,
    when a reference tree is reachable ($EGOTAP_REF, /root/reference, baseline/_ref); None otherwise (the GPU box)."""
    import ref_shim
    if ref_shim.reference_root() is None:
        return None
    import torch
    from egotap_b200.synthetic import synthetic_heatmaps
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    net = ref_shim.build_reference_net(preset)
    net.load_state_dict(sd, strict=True)
    x = synthetic_heatmaps(preset, batch, seed=1234, kind="gauss")
    with torch.no_grad():
        for _ in range(warmup):
            net(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            pose = net(x)[0].detach()
        dt = (time.perf_counter() - t0) / steps
    return dict(fps=batch / dt, ms_per_step=dt * 1e3, cores=torch.get_num_threads(), batch=batch, pose=pose, x=x)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    import weights
    sd = weights.make_state_dict(args.preset, seed=0, randomize=False)
    warm = max(1, min(args.warmup, 2))
    kind = "reference"
    r = time_reference_module_cpu(args.preset, sd, steps=args.steps, warmup=warm)
    if r is None:            # no reference tree here: the oracle port (pinned to the reference to ~1e-6, the same torch ops)
        kind = "port"
        r = time_oracle_cpu(args.preset, sd, steps=args.steps, warmup=warm)
    sample = "%d steps x batch %d frames, fp32, %d host threads" % (args.steps, r["batch"], r["cores"])
    line = dict(impl="reference", metric=METRIC, value=r["fps"], unit="frames/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=workload_name(args), preset=args.preset, batch_per_step=r["batch"]),
                cpu_baseline=dict(value=r["fps"], unit="frames/s", cores=r["cores"], kind=kind, sample=sample),
                e2e=dict(value=r["fps"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    print(json.dumps(line), flush=True)


def run_torch_eager(args):
    """Context arm (not part of the driver contract): the SAME algorithm through stock PyTorch eager kernels
    (cuBLAS / ATen) on the GPU -- the 'kernel to beat on the same box', since the reference ships no CUDA of its
    own.  Uses the oracle restatement, fp32 (TF32 off) and bf16 autocast."""
    import torch
    import egotap_oracle as orc
    import weights
    from egotap_b200.synthetic import synthetic_heatmaps
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = {k: v.to(dev) for k, v in weights.make_state_dict(args.preset, seed=0, randomize=False).items()}
    x = synthetic_heatmaps(args.preset, args.batch, seed=1234, kind="gauss").to(dev)
    out = {}
    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16_autocast", torch.autocast("cuda", dtype=torch.bfloat16))):
        with torch.no_grad(), ctx:
            for _ in range(max(1, args.warmup)):
                orc.forward(sd, x, args.preset)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                orc.forward(sd, x, args.preset)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[name] = dict(value=args.batch / ms * 1e3, unit="frames/s", ms_per_step=ms)
    print(json.dumps(dict(impl="torch_eager", metric=METRIC, n_gpus=1, steps=args.steps, warmup=args.warmup,
                          config=dict(workload=workload_name(args), preset=args.preset, batch_per_gpu=args.batch),
                          value=out["fp32"]["value"], unit="frames/s", dtype="f32 (TF32 off)", arms=out)), flush=True)


def workload_name(args):
    if getattr(args, "workload", "lifting") == "lifting_gt":
        return "2-D/3-D keypoints -> ground-truth joint+limb heatmaps synthesised on the GPU (gt_heatmap_kernel) -> EgoTAP " \
               "lifting net, %s preset, random-init weights, batch %d per GPU (the reference's --use_gt_heatmap path)" \
               % (args.preset, args.batch)
    if getattr(args, "workload", "lifting") == "e2e_rgb":
        return "stereo RGB 256x256 -> 2 x ResNet-18 U-Net heatmap producers (torch/cuDNN, bf16 autocast) -> EgoTAP " \
               "lifting net (sm_100a kernels), %s preset, random-init weights, batch %d per GPU" % (args.preset, args.batch)
    return "EgoTAP lifting net (3-layer ViT heatmap encoder + limb FC encoder + 2-layer propagation chain + head), " \
           "%s preset, random-init weights, synthetic stereo joint+limb heatmaps 64x64, batch %d per GPU" % (args.preset, args.batch)


class Ctx:
    """process-wide state of one bench.py run: rank / device / process group"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if args.gpus != self.world and self.world > 1:
            raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, self.world))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()


PARITY_FRAMES = 16


def measure_lifting(ctx, args, preset, B, precision, workload="lifting", graph=False, cpu_baseline=False, dump=""):
    """One timed measurement of the lifting path on this process group: W warm-up steps, parity of what is being timed
    (rank 0, PARITY_FRAMES frames through the CPU oracle), K timed steps with inputs resident in HBM, K timed steps end to end
    from pinned host memory, then (rank 0) the per-kernel rooflines of one extra step.  Returns the result dict on rank 0,
    None elsewhere.  Every rank runs its own shard of B frames per step; the path's only collective is ONE gather of all
    poses of the job at its end ("a final gather"), inside the timed region."""
    torch, dist = ctx.torch, ctx.dist
    rank, world, dev, local_rank = ctx.rank, ctx.world, ctx.dev, ctx.local_rank
    import egotap_b200
    import egotap_oracle as orc
    from egotap_b200 import capi
    from egotap_b200.pipeline import HostPipeline
    from egotap_b200.sharded import gather_job_poses
    from egotap_b200.options import make_opt

    K, W = args.steps, args.warmup
    torch.manual_seed(0)                                   # reference-style random init (kaiming), same on every rank
    net = egotap_b200.EgoTAPAutoEncoder(make_opt(preset, b200_precision=precision, b200_max_batch=B,
                                                 b200_cuda_graph=B if (graph and workload != "e2e_rgb") else 0), input_channel_scale=2)
    net.init_weights("kaiming")
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.to(dev).eval()
    x_host = egotap_b200.synthetic_heatmaps(preset, B, seed=1234 + rank, kind="gauss").pin_memory()
    x = x_host.to(dev)
    total = B * world
    est = None
    if workload == "e2e_rgb":
        # BASELINE config 4: torch/cuDNN ResNet-18 U-Net heatmap producers (random init) feeding the lifting kernels
        from egotap_b200.heatmap_net import HeatMapUNet, StereoPoseEstimator
        pos_opt, rot_opt = make_opt(preset), make_opt(preset)
        pos_opt.num_rot_heatmap = 0
        rot_opt.num_heatmap = 0
        est = StereoPoseEstimator(HeatMapUNet(pos_opt).to(dev).to(memory_format=torch.channels_last),
                                  HeatMapUNet(rot_opt).to(dev).to(memory_format=torch.channels_last), net,
                                  cuda_graph=bool(graph)).eval()
        g = torch.Generator().manual_seed(77 + rank)
        rgb_host = [torch.rand(B, 3, 256, 256, generator=g).pin_memory() for _ in range(2)]
        rgb = [t.to(dev) for t in rgb_host]

    kp = None
    if workload == "lifting_gt":
        # the reference's --use_gt_heatmap path: heatmaps synthesised from keypoints (here on the GPU, per step)
        import numpy as np
        from egotap_b200.gt_heatmaps import synthesize
        n_pts = 16 if preset == "UnrealEgo" else 18
        rng = np.random.default_rng(1234 + rank)
        kp_host = (torch.from_numpy(rng.uniform(0, 1024, size=(B, 2, n_pts, 2)).astype("float32")).pin_memory(),
                   torch.from_numpy(rng.normal(0, 30, size=(B, n_pts, 3)).astype("float32")).pin_memory())
        kp = [t.to(dev) for t in kp_host]
        x = synthesize(kp[0], kp[1], preset)
        x_host = x.cpu()

    def local_step():          # this rank's shard only: no collective
        if kp is not None:
            return net.predict_pose(synthesize(kp[0], kp[1], preset, out=x))
        return est(rgb[0], rgb[1]) if est is not None else net.predict_pose(x)

    for _ in range(W):
        pose = local_step()
    if world > 1:                 # the communicator's first all-gather sets up its channels (~100 ms): not part of the job
        gather_job_poses(torch.stack([pose, pose]))
    torch.cuda.synchronize()
    # ---------------- parity of what is being timed (PARITY_FRAMES frames through the oracle, rank 0)
    parity = None
    if rank == 0:
        # torchrun exports OMP_NUM_THREADS=1; the other ranks wait at the barrier below while this check runs
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0)) // world))
        n_par = min(PARITY_FRAMES, B)
        with torch.no_grad():
            ref_in = est.pred_heatmap_cat[:n_par].float().cpu() if est is not None else x_host[:n_par]
            ref = orc.forward(sd, ref_in, preset)
        parity = orc.parity_report(pose[:n_par], ref)
        parity["frames"] = n_par
    # ---------------- timed region: inputs resident in HBM; K steps per rank, then the job's one gather
    launches0 = capi.lib().egotap_b200_launch_count()
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    el = torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        e0.record()
        poses = [local_step() for _ in range(K)]
        el.record()                                         # this rank's own compute is done here
        if world > 1:
            gathered = gather_job_poses(torch.stack(poses))
        e1.record()
        torch.cuda.synchronize()
    ms, ms_local = e0.elapsed_time(e1), e0.elapsed_time(el)
    launches = capi.lib().egotap_b200_launch_count() - launches0
    if world > 1:
        assert tuple(gathered.shape) == (world, K) + tuple(poses[0].shape)
    # ---------------- end to end: pinned host buffers in, pinned host poses out, copies inside the timed region
    host_out = [torch.empty((B, net.num_joints, 3)).pin_memory() for _ in range(K)]
    if kp is not None:
        h2d, d2h = kp_host[0].numel() * 4 + kp_host[1].numel() * 4, B * net.num_joints * 3 * 4

        def run_e2e():
            for i in range(K):
                a = kp_host[0].to(dev, non_blocking=True)
                b = kp_host[1].to(dev, non_blocking=True)
                host_out[i].copy_(net.predict_pose(synthesize(a, b, preset, out=x)), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        run_e2e()
    elif est is None:
        pipe = HostPipeline(net, B)
        host_in = [x_host, x_host.clone().pin_memory()]
        pipe.run([host_in[i % 2] for i in range(min(W, 2))], host_out)
        h2d, d2h = pipe.h2d_bytes_per_step, pipe.d2h_bytes_per_step

        def run_e2e():
            pipe.run([host_in[i % 2] for i in range(K)], host_out)
    else:
        h2d, d2h = 2 * rgb_host[0].numel() * 4, B * net.num_joints * 3 * 4

        def run_e2e():
            for i in range(K):
                l = rgb_host[0].to(dev, non_blocking=True)
                r = rgb_host[1].to(dev, non_blocking=True)
                host_out[i].copy_(est(l, r), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        run_e2e()
    ctx.barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    run_e2e()
    t1.record()
    torch.cuda.synchronize()
    ms_e2e = t0.elapsed_time(t1)
    rank_ms = [ms_local]
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_local], device=dev)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        allt = torch.stack(allt).cpu()
        ms, ms_e2e = float(allt[:, 0].max()), float(allt[:, 1].max())
        rank_ms = allt[:, 2].tolist()
    if rank != 0:
        ctx.barrier()            # rank 0 still has collective-free work (per-kernel profile, CPU baseline) before the next region
        return None
    e2e_ok = bool(torch.isfinite(host_out[-1]).all()) and (host_out[-1][:2].to(dev) - pose[:2]).abs().max().item() < 1e-5
    # ---------------- roofline of the dominant kernel family (tcgen05 GEMM), per-launch CUDA events, one extra step
    pk = peaks()
    # rank 0 alone runs this extra step (the other ranks wait at the barrier): it must not contain a collective
    graph_batch, net._graph_max_batch = net._graph_max_batch, 0    # the per-kernel events need launched (not replayed) kernels
    if est is not None:
        est_graph, est.cuda_graph = est.cuda_graph, False
    all_recs = capi.profile_kernels(lambda: local_step())
    net._graph_max_batch = graph_batch
    if est is not None:
        est.cuda_graph = est_graph
    if graph:
        launches = K * len(all_recs)                                # kernels executed from the graph in the timed region
    recs = [r for r in all_recs if "flops" in r]
    gemm_ms = sum(r["ms"] for r in recs)
    gemm_flops = sum(r["flops"] for r in recs)
    if not recs or gemm_ms <= 0:
        raise SystemExit("bench: the per-kernel profile of one step recorded no GEMM launch -- the CUDA path did not run")
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
    nsplit = 3 if precision != "bf16" else 1
    # DRAM traffic of the same launches from the committed ncu capture (profiles/), when it is for this configuration
    traffic, traffic_src = None, None
    for name in ("r02_traffic_%s.json" % precision, "r01b_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            if (tj["batch"], tj["preset"], tj["precision"]) == (B, preset, precision):
                traffic, traffic_src = tj["gemm_tc_kernel_bytes_per_step"], "profiles/" + name
                break
    step_ms = ms_local / K
    roofline = dict(bound="tensor", kernel="gemm_tc_kernel (all %d launches of one step)" % len(recs), achieved=achieved,
                    peak=pk["bf16_sustained"], unit="TFLOP/s", frac=achieved / pk["bf16_sustained"],
                    frac_vs_burst=achieved / pk["bf16_burst"], traffic=traffic,
                    traffic_note="DRAM bytes of the same GEMM launches of one step (ncu, %s)" % traffic_src,
                    peak_source=pk["source"] + ", sustained bf16 figure (kernel timed inside a long step)",
                    mma_passes_per_flop=nsplit, mma_frac=achieved * nsplit / pk["bf16_sustained"],
                    gemm_share_of_step=gemm_ms / step_ms, note="achieved = algorithmic 2*M*N*K over all GEMM launches of one "
                    "step / their summed CUDA-event durations; bf16x3 issues 3 MMAs per algorithmic FLOP")
    # HBM-bound kernels: algorithmic bytes per launch / CUDA-event duration vs measured copy peak
    J = 15 if preset == "UnrealEgo" else 17
    nb = 2 * (2 if nsplit == 3 else 1)                       # bytes per operand element written (bf16 hi [+ lo])
    rows_ln = B * 576
    hbm_bytes = {"layernorm1024_kernel": rows_ln * 1024 * (4 + nb),
                 "ingest_kernel": B * 6 * J * 4096 * (4 + nb),
                 "gt_heatmap_kernel": B * 6 * J * 4096 * 4}              # the heatmap stack written once; inputs are ~0.5 KB/frame
    hbm = {}
    for name, nbytes in hbm_bytes.items():
        ts = [r["ms"] for r in all_recs if r["name"] == name]
        if ts:
            t = max(ts) if name == "layernorm1024_kernel" else ts[0]   # full-size LN launches (the last one is compacted)
            hbm[name] = dict(ms=t, bytes=nbytes, achieved_gbs=nbytes / t / 1e6, peak_gbs=pk["hbm_gbs"],
                             frac=nbytes / t / 1e6 / pk["hbm_gbs"])
    # fused attention (second-largest kernel family): algorithmic 4*T*T*d FLOP per (frame, head) over the query rows each
    # layer needs (all 576 in layers 0-1, the live tokens in the last one), CUDA-event durations of the same step
    att = [r["ms"] for r in all_recs if r["name"].startswith("attention")]
    attention = None
    if att:
        att_flops = 4.0 * 128 * 576 * (576 + 576 + 2 * J * 16) * 8 * B
        att_tf = att_flops / (sum(att) * 1e-3) / 1e12
        attention = dict(kernel=sorted({r["name"] for r in all_recs if r["name"].startswith("attention")})[0], launches=len(att),
                         ms=sum(att), achieved=att_tf, peak=pk["bf16_sustained"], unit="TFLOP/s", frac=att_tf / pk["bf16_sustained"],
                         mma_frac=att_tf * nsplit / pk["bf16_sustained"], share_of_step=sum(att) / step_ms)
    fps = total * K / (ms * 1e-3)
    flop_frame = FLOP_PER_FRAME[preset] + (107.41e9 if est is not None else 0.0)   # + producers
    whole = dict(achieved=fps / world * flop_frame / 1e12, peak=pk["bf16_sustained"], unit="TFLOP/s")
    whole["frac"] = whole["achieved"] / whole["peak"]
    whole["frac_vs_burst"] = whole["achieved"] / pk["bf16_burst"]
    # ---------------- CPU baseline on this box's host cores (bounded sample)
    cpu = time_oracle_cpu(preset, sd, steps=3, warmup=1) if (world == 1 and cpu_baseline) else None
    wl = argparse.Namespace(workload=workload, preset=preset, batch=B)
    res = dict(value=fps, unit="frames/s", ms_per_step=ms / K,
               dtype="bf16x3 operands, f32 accumulate" if nsplit == 3 else "bf16 operands, f32 accumulate",
               config=dict(workload=workload_name(wl), preset=preset, batch_per_gpu=B, global_batch=total,
                           precision=precision, issue="cuda graph replay" if graph else "%d launches per step" % len(all_recs),
                           switches={k: v for k, v in sorted(os.environ.items()) if k.startswith("EGOTAP_")},   # A/B switches in effect
                           parallelism="dp%d (frames sharded, no data-path collective, one gather of all poses at the end of "
                                       "the job inside the timed region)" % world,
                           l2="inputs %.0f MB + activations >> 126 MB L2 per step, no flush needed" % (x.numel() * 4 / 1e6)),
               e2e=dict(value=total * K / (ms_e2e * 1e-3), unit="frames/s", h2d_bytes_per_step=h2d,
                        d2h_bytes_per_step=d2h, checked=e2e_ok),
               gpu_launches=int(launches), clocks=clocks.summary(),
               rank_ms=dict(min=min(rank_ms) / K, max=max(rank_ms) / K, job_ms_per_step=ms / K,
                            note="per-rank CUDA-event time of the rank's own K steps / K (before the final gather); the job time "
                                 "is the max over ranks incl. the gather"),
               roofline=roofline, roofline_whole_step=whole,
               roofline_hbm_kernels=hbm, roofline_attention=attention,
               cpu_baseline=(dict(value=cpu["fps"], unit="frames/s", cores=cpu["cores"], kind="port",
                                  sample="3 steps x batch %d frames of the same workload, fp32 oracle" % cpu["batch"])
                             if cpu else None),
               parity=parity)
    if dump:
        os.makedirs(os.path.dirname(os.path.abspath(dump)), exist_ok=True)
        with open(dump, "w") as f:
            json.dump(dict(line=res, kernel_launches=all_recs), f, indent=1)
    del net, est
    torch.cuda.empty_cache()
    ctx.barrier()
    return res


CONFIG3_GLOBAL_BATCH = 1024


def run_ours(args):
    """The driver's line: BASELINE config 2 (UnrealEgo, 256 frames per GPU, fp32-parity mode, weak scaling) as the headline,
    and in the same run (unless --only-headline or a non-default configuration was asked for): ``modes.bf16`` = the same
    workload in the bf16-operand throughput mode (its own stated parity bound), and ``config3`` = BASELINE config 3 (EgoCap
    preset, 1024 frames GLOBAL, bf16 operands, sharded over the N GPUs: strong scaling)."""
    ctx = Ctx(args)
    default_cfg = (args.workload, args.preset, args.batch, args.precision, args.graph) == ("lifting", "UnrealEgo", 256, "bf16x3", False)
    main = measure_lifting(ctx, args, args.preset, args.batch, args.precision, workload=args.workload, graph=args.graph,
                           cpu_baseline=True, dump=args.dump)
    extra = {}
    if default_cfg and not args.only_headline:
        extra["modes"] = dict(bf16=measure_lifting(ctx, args, "UnrealEgo", 256, "bf16"))
        b3 = CONFIG3_GLOBAL_BATCH // ctx.world
        c3 = measure_lifting(ctx, args, "EgoCap", b3, "bf16")
        if c3 is not None:
            c3["scaling"] = "strong"
        extra["config3"] = c3
    if ctx.rank == 0:
        line = dict(metric=METRIC, value=main["value"], unit=main["unit"], n_gpus=ctx.world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=main["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype=main["dtype"],
                    data="synthetic")
        line.update({k: v for k, v in main.items() if k not in line})
        line.update(extra)
        print(json.dumps(line), flush=True)
    _leave(ctx.dist, ctx.world)


def _leave(dist, world):
    """symmetric teardown: rank 0 still has collective-free work to do (per-kernel profile, JSON line) after the other ranks
    are finished; everybody meets once more before the process group is destroyed, so that no rank tears its communicator
    down while a peer is still alive in it"""
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def time_train_oracle_cpu(preset, steps, batch=4):
    """The reference's optimisation step (train-mode forward, MPJPE + bone-cosine loss, autograd backward, AdamW) as
    restated by oracle/train_oracle.py, fp32, all host threads, bounded sample."""
    import torch
    import train_oracle as tro
    import weights
    from egotap_b200.synthetic import synthetic_heatmaps
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    sd = weights.make_state_dict(preset, seed=0, randomize=False)
    x = synthetic_heatmaps(preset, batch, seed=1234, kind="gauss")
    nj = 16 if preset == "UnrealEgo" else 17
    gt = torch.randn(batch, nj, 3, generator=torch.Generator().manual_seed(7)) * 20
    _, sd, state, _ = tro.train_step(sd, x, gt, preset)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss, sd, state, _ = tro.train_step(sd, x, gt, preset, opt_state=state)
    dt = (time.perf_counter() - t0) / steps
    return dict(fps=batch / dt, ms_per_step=dt * 1e3, cores=torch.get_num_threads(), batch=batch, loss=float(loss))


TRAIN_METRIC = "train_frames_per_sec_lifting_net"
TRAIN_WORKLOAD = "EgoTAP pose-estimator training step (train-mode forward, MPJPE + bone-cosine loss, backward, AdamW), " \
                 "%s preset, random-init weights, synthetic heatmaps + poses, batch %d per GPU"


def run_train_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    r = time_train_oracle_cpu(args.preset, steps=max(1, min(args.steps, 3)))
    sample = "%d steps x batch %d frames, fp32, %d host threads" % (max(1, min(args.steps, 3)), r["batch"], r["cores"])
    print(json.dumps(dict(impl="reference", metric=TRAIN_METRIC, value=r["fps"], unit="frames/s", n_gpus=args.gpus,
                          steps=args.steps, warmup=args.warmup, ms_per_step=r["ms_per_step"], higher_is_better=True,
                          scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                          config=dict(workload=TRAIN_WORKLOAD % (args.preset, r["batch"]), preset=args.preset),
                          cpu_baseline=dict(value=r["fps"], unit="frames/s", cores=r["cores"], kind="port", sample=sample),
                          e2e=dict(value=r["fps"], unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                          gpu_launches=0)), flush=True)


def run_train(args):
    """BASELINE config 5: one optimisation step per 'step', per-GPU micro-batch fixed (weak scaling), gradients
    all-reduced over NCCL in backward-completion order while the backward is still running."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import egotap_b200
    from egotap_b200 import capi, ddp
    from egotap_b200.options import make_opt
    B, K, W = args.batch, args.steps, args.warmup
    torch.manual_seed(0)
    net = egotap_b200.EgoTAPAutoEncoder(make_opt(args.preset, b200_precision=args.precision), input_channel_scale=2)
    net.init_weights("kaiming")
    net = net.to(dev).train()
    eng = net.train_engine()
    eng.use_cuda_graph = bool(args.graph) and world == 1
    eng.persistent_bptt = bool(args.persistent_bptt)
    red = ddp.StagedGradAllReduce(eng) if world > 1 else None
    nj = net.num_joints
    x_host = egotap_b200.synthetic_heatmaps(args.preset, B, seed=1234 + rank, kind="gauss").pin_memory()
    gt_host = (torch.randn(B, nj, 3, generator=torch.Generator().manual_seed(7 + rank)) * 20).pin_memory()
    x, gt = x_host.to(dev), gt_host.to(dev)
    losses = []
    for _ in range(W):
        losses.append(eng.train_step(x, gt, reducer=red).clone())
    torch.cuda.synchronize()
    launches0 = capi.lib().egotap_b200_launch_count()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if red is not None:
        red.time_exposed = True
    with ClockSampler(local_rank) as clocks:
        e0.record()
        for _ in range(K):
            loss = eng.train_step(x, gt, reducer=red)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    comm = None
    if red is not None:
        comm = dict(exposed_ms_per_step=red.exposed_ms(), bytes_per_step=int(eng.flat_grad.numel()) * 4, collectives_per_step=len(red.launched),
                    note="gradient all-reduce (NCCL, SUM) launched per >= 32 MB slice of the flat gradient buffer as soon as the backward "
                         "has finished it; exposed = CUDA-event time the compute stream waits for the collectives before AdamW (rank 0)")
        red.time_exposed = False
    launches = capi.lib().egotap_b200_launch_count() - launches0
    losses.append(loss.clone())
    # end to end: inputs and targets from pinned host memory every step, the loss read back every step
    loss_host = torch.empty(4).pin_memory()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    # the next batch is uploaded on a copy stream into one of two staging buffers while the current step runs (the engine
    # copies its inputs into its own buffers at the start of a step, so a staging buffer is free again after that step)
    copy_stream = torch.cuda.Stream(device=dev)
    stage = [(torch.empty_like(x), torch.empty_like(gt)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    compute = torch.cuda.current_stream(dev)

    def upload(i):
        s_ = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[s_])
            stage[s_][0].copy_(x_host, non_blocking=True)
            stage[s_][1].copy_(gt_host, non_blocking=True)
            ready[s_].record(copy_stream)
    for ev in done:
        ev.record(compute)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    upload(0)
    for i in range(K):
        if i + 1 < K:
            upload(i + 1)
        compute.wait_event(ready[i % 2])
        loss_host.copy_(eng.train_step(stage[i % 2][0], stage[i % 2][1], reducer=red), non_blocking=True)
        done[i % 2].record(compute)
    t1.record()
    torch.cuda.synchronize()
    ms_e2e = t0.elapsed_time(t1)
    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        _leave(dist, world)
        return
    pk = peaks()
    all_recs = capi.profile_kernels(lambda: (eng.forward(x), eng.loss_and_grad(gt), eng.backward(), eng.adamw_step(lr=0.0)))
    recs = [r for r in all_recs if "flops" in r]
    gemm_ms, gemm_flops = sum(r["ms"] for r in recs), sum(r["flops"] for r in recs)
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
    nsplit = 3 if args.precision != "bf16" else 1
    by_kernel = {}
    for r in all_recs:
        by_kernel[r["name"]] = by_kernel.get(r["name"], 0.0) + r["ms"]
    total = B * world
    fps = total * K / (ms * 1e-3)
    cpu = time_train_oracle_cpu(args.preset, steps=2) if world == 1 else None
    line = dict(metric=TRAIN_METRIC, value=fps, unit="frames/s", n_gpus=world, steps=K, warmup=W, ms_per_step=ms / K,
                higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="bf16x3 operands, f32 accumulate" if nsplit == 3 else "bf16 operands, f32 accumulate, f32 master weights",
                data="synthetic",
                config=dict(workload=TRAIN_WORKLOAD % (args.preset, B), preset=args.preset, batch_per_gpu=B, global_batch=total,
                            precision=args.precision, issue="cuda graph" if eng.use_cuda_graph else "recorded-call replay",
                            chain_backward="persistent kernel" if eng.persistent_bptt else "per-joint launches",
                            parallelism="dp%d (per-rank micro-batch, staged gradient all-reduce over NCCL)" % world,
                            l2="activations + gradients >> 126 MB L2 per step, no flush needed"),
                e2e=dict(value=total * K / (ms_e2e * 1e-3), unit="frames/s", h2d_bytes_per_step=x_host.numel() * 4 + gt_host.numel() * 4,
                         d2h_bytes_per_step=16, checked=bool(torch.isfinite(loss_host).all())),
                gpu_launches=int(launches), clocks=clocks.summary(),
                roofline=dict(bound="tensor", kernel="gemm_tc_kernel (all %d launches of one training step)" % len(recs),
                              achieved=achieved, peak=pk["bf16_sustained"], unit="TFLOP/s", frac=achieved / pk["bf16_sustained"],
                              traffic=None, mma_passes_per_flop=nsplit, gemm_share_of_step=gemm_ms / (ms / K),
                              peak_source=pk["source"] + ", sustained bf16 figure"),
                kernel_ms_per_step={k: round(v, 4) for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1])},
                allreduce=comm,
                loss_trace=[float(l[0]) for l in losses],
                cpu_baseline=(dict(value=cpu["fps"], unit="frames/s", cores=cpu["cores"], kind="port",
                                   sample="2 steps x batch %d frames of the same workload, fp32 autograd oracle" % cpu["batch"])
                              if cpu else None))
    print(json.dumps(line), flush=True)
    if args.dump:
        os.makedirs(os.path.dirname(os.path.abspath(args.dump)), exist_ok=True)
        with open(args.dump, "w") as f:
            json.dump(dict(line=line, kernel_launches=all_recs), f, indent=1)
    _leave(dist, world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="egotap_b200", choices=["egotap_b200", "reference", "torch_eager"])
    ap.add_argument("--preset", default="UnrealEgo", choices=["UnrealEgo", "EgoCap"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: 256 for the lifting workloads, 32 for train)")
    ap.add_argument("--precision", default=None, choices=["bf16x3", "bf16"],
                    help="operand precision (default: bf16x3 = fp32-parity mode for the lifting workloads, bf16 for train)")
    ap.add_argument("--workload", default="lifting", choices=["lifting", "lifting_gt", "e2e_rgb", "train"],
                    help="lifting = BASELINE configs 1-3 (default); lifting_gt = the same with the input heatmaps synthesised on "
                         "the GPU from keypoints each step (--use_gt_heatmap path); e2e_rgb = config 4 (RGB -> heatmap nets -> lifting); "
                         "train = config 5 (optimisation step; reference batch size 32, scripts/train/PoseEstimator/*.sh)")
    ap.add_argument("--persistent-bptt", dest="persistent_bptt", action="store_true",
                    help="train workload: BPTT of each propagation layer as one persistent launch instead of per-joint launches")
    ap.add_argument("--graph", action="store_true", help="train workload, 1 GPU: run forward+loss+backward from a CUDA graph; "
                    "lifting workloads: replay the forward from a CUDA graph (opt b200_cuda_graph; small-batch serving)")
    ap.add_argument("--only-headline", dest="only_headline", action="store_true",
                    help="lifting workload: skip the extra timed regions (modes.bf16, config3) of the default run")
    ap.add_argument("--dump", default="", help="also write the JSON line + per-GEMM launch table to this file")
    args = ap.parse_args()
    if args.batch <= 0:
        args.batch = 32 if args.workload == "train" else 256
    if args.precision is None:
        args.precision = "bf16" if args.workload == "train" else "bf16x3"
    if args.warmup < 3 and args.impl != "reference":
        args.warmup = 3
    if args.workload == "train":
        if args.impl == "reference":
            run_train_reference(args)
        elif args.impl == "egotap_b200":
            if args.gpus > 1 and "RANK" not in os.environ:
                raise SystemExit("launch N > 1 under torch.distributed.run (see the module docstring)")
            run_train(args)
        else:
            raise SystemExit("--workload train has no torch_eager arm")
    elif args.impl == "reference":
        run_reference(args)
    elif args.impl == "torch_eager":
        run_torch_eager(args)
    else:
        if args.gpus > 1 and "RANK" not in os.environ:
            raise SystemExit("launch N > 1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N "
                             "--master-addr 127.0.0.1 --master-port P bench.py --gpus N ...")
        run_ours(args)


if __name__ == "__main__":
    main()
