#!/usr/bin/env python
"""Pre-training throughput of the GeoMAE hot path on B200 (frames/sec), one JSON line.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # CPU port of the reference path (oracle/), rank 0 only

A "step" is one full training step of the mae_sst nuScenes config on `samples_per_gpu` synthetic
nuScenes-shaped frames per GPU: 3-scale voxelise+scatter -> geometric targets -> VFE -> SRA encoder /
decoders -> 6 losses -> backward -> gradient all-reduce -> grad-clip + AdamW.
`value` times K steps with the frames already resident in HBM; `e2e` times the same K steps through
FlatTrainer.train_step_from_host (pinned host frames, H2D inside the timed region, loss read back).
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OWN_CFG = os.path.join(ROOT, "configs/mae_sst/geomae_nus_pretrain.py")
METRIC = "pretrain frames/sec on nuScenes-shaped synthetic sweeps"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="nus", choices=["nus", "nus10sweep", "waymo", "dense"],
                    help="nus = BASELINE.json configs[1] (the metric's configuration); the others are configs[3]/[4] and the "
                         "10-sweep nuScenes input (geomae_b200/workloads.py)")
    ap.add_argument("--samples-per-gpu", type=int, default=0)     # 0: the workload's default (4 = data.samples_per_gpu)
    ap.add_argument("--sweeps", type=int, default=0)              # 0: the workload's default
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--length-bins", action="store_true", help="add the SRA window-length sweep (default for --workload dense)")
    ap.add_argument("--list-only", action="store_true",
                    help="profiler runs: W warm-up + K resident steps, then exit without the e2e / roofline / CPU legs")
    ap.add_argument("--sra-impl", default="tc1", choices=["tc1", "tc3", "glue"],
                    help="tc1: bf16 tensor-core SRA layers (BASELINE config 'bf16'); tc3: bf16x3 split (fp32 parity); "
                         "glue: library GEMMs")
    return ap.parse_args()


def workload_of(args):
    from geomae_b200.workloads import WORKLOADS
    w = dict(WORKLOADS[args.workload])
    w["frame"] = dict(w["frame"])
    if args.sweeps:
        w["frame"]["sweeps"] = args.sweeps
    if args.samples_per_gpu:
        w["samples_per_gpu"] = args.samples_per_gpu
    return w


def make_batches(rank, n_batches, samples, frame_kw):
    from geomae_b200.synthetic import make_frame
    return [[make_frame(1000 * rank + it * 16 + s + 1, **frame_kw) for s in range(samples)] for it in range(n_batches)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.windows = index, [], None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")] + [time.time()])

    def wait_ready(self, timeout=5.0):
        """Block until the first sample arrived, so nvidia-smi's initialisation is over before anything is timed."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """Wall-clock bracket of a timed region: only samples taken inside brackets are reported."""
        self.windows.append(time.time())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        w = self.windows
        inside = lambda t: any(w[i] <= t <= w[i + 1] for i in range(0, len(w) - 1, 2)) if len(w) >= 2 else True  # noqa: E731
        rows = [r for r in self.rows if len(r) >= 10 and inside(r[-1])]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


def hbm_kernel_rooflines(model, pool, dev, hbm_peak, peak_src, n_frames=None):
    """HBM-bound stages on a batch large enough to leave L2 (~7 M points, 140 MB of records: 256 one-sweep frames):
    achieved = algorithmic bytes (SURVEY.md §8d formulas, counted from the device-side totals) / CUDA-event time."""
    from geomae_b200.voxel import scatter_frames
    if n_frames is None:
        n_frames = max(8, min(256, int(round(7.0e6 / max(1, pool[0][0].shape[0])))))
    frames = [torch.from_numpy(pool[i % len(pool)][i % len(pool[0])]).to(dev) for i in range(n_frames)]
    pb = scatter_frames(model.geom, frames)
    v, vm, vl = pb.sizes()
    p = pb.points.shape[0]

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t_sc = timed(pb.run)
    t_gt = timed(lambda: pb.geom_targets())
    b_sc = 24.0 * p + 32.0 * v + 20.0 * (vm + vl)
    b_gt = 20.0 * vm + 28.0 * v + 24.0 * v
    # the data step in front of the scatter: rotate / scale / flip / range-filter / compact, two
    # passes over the 20-byte records: read 12 B (xyz) to count, read 20 B + write 20 B to move = 52 B per raw point
    from geomae_b200.data import augment_filter, draw_augmentation
    rs = np.random.RandomState(7)
    augs = [draw_augmentation(rs) for _ in range(n_frames)]
    t_au = timed(lambda: augment_filter(pb.points, pb.frame_offsets, augs, model.point_cloud_range))
    b_au = 52.0 * p
    # DRAM traffic of the same stages from the committed ncu --set full capture (same batch size, 256 frames)
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r01_v4_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("frames") == n_frames:
            traffic = {"scatter": tj["voxel_scatter"]["traffic_bytes"], "geom": tj["k_geom"]["traffic_bytes"],
                       "source": "profiles/r01_v4_traffic.json"}
    except (OSError, ValueError, KeyError):
        pass
    out = []
    for name, key, t, b in (("geomae_voxel_scatter (9 kernels)", "scatter", t_sc, b_sc),
                            ("k_geom (geom_targets)", "geom", t_gt, b_gt),
                            ("geomae_augment_filter (3 kernels; data step N2)", "augment", t_au, b_au)):
        ach = b / (t * 1e-3) / 1e9
        out.append(dict(kernel=name, bound="hbm", frames=n_frames, points=p, pillars=v, ms=t, algorithmic_bytes=b,
                        achieved=ach, peak=hbm_peak, unit="GB/s", frac=ach / hbm_peak, peak_source=peak_src,
                        traffic=traffic.get(key), traffic_source=traffic.get("source")))
    return out


def sra_length_bin_sweep(dev, tens_peak, tokens=196608):
    """BASELINE.json configs[4] "SRA length-bin sweep": the window-attention kernels on synthetic CSR layouts whose
    windows ALL have length L (the reference's bucket sizes), ~200 k tokens each; flops = 64 / 160 per (query, key, head)."""
    from geomae_b200 import lib as L
    out = []
    for Lw in (16, 32, 64, 144):
        nw = tokens // Lw
        n = nw * Lw
        win_ptr = torch.arange(0, n + 1, Lw, dtype=torch.int32, device=dev)
        g = torch.Generator(device="cpu").manual_seed(Lw)
        perm = torch.randperm(n, generator=g).to(dev)
        win_tok = perm.int()                                   # window w holds tokens perm[w*L : (w+1)*L]
        tok_win = torch.empty(n, dtype=torch.int32, device=dev)
        tok_win[perm] = (torch.arange(n, device=dev) // Lw).int()
        qkv = (torch.randn(n, 384, device=dev) * 0.5).to(torch.bfloat16)
        attn = torch.empty(n, 128, dtype=torch.bfloat16, device=dev)
        lse = torch.empty(n, 8, dtype=torch.float32, device=dev)
        d_out = (torch.randn(n, 128, device=dev) * 0.5).to(torch.bfloat16)
        dd = torch.randn(n, 8, device=dev)
        dqkv = torch.empty_like(qkv)
        st = L.stream_ptr(dev)

        def fwd():
            L.run("sra_attention_tc_fwd", L.ptr(qkv), n, 8, L.ptr(win_ptr), L.ptr(win_tok), L.ptr(tok_win), L.ptr(attn),
                  L.ptr(lse), 1 | 8, st)

        def bwd():
            L.run("sra_attention_tc_bwd", L.ptr(qkv), L.ptr(attn), L.ptr(lse), L.ptr(d_out), n, 8, L.ptr(win_ptr),
                  L.ptr(win_tok), L.ptr(tok_win), L.ptr(dqkv), L.ptr(dd), 1 | 2 | 4, st)
        res = dict(window_length=Lw, windows=nw, tokens=n)
        for name, fn, fl in (("fwd", fwd, 64.0), ("bwd", bwd, 160.0)):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms_ = e0.elapsed_time(e1) / 5
            tf = fl * 8 * nw * Lw * Lw / (ms_ * 1e-3) / 1e12
            res[name] = dict(ms=ms_, tflops=tf, tensor_frac=tf / tens_peak, tokens_per_us=n / (ms_ * 1e3))
        out.append(res)
    return out


def cpu_reference_step(samples, frame_kw, threads, seed=1, geometry=None, blocks=None):
    """One forward+backward of the oracle port on the host cores; returns (seconds, frames).  blocks=(enc, dec)
    shrinks the model (BASELINE.json configs[0]: "1 SST block")."""
    from geomae_b200.synthetic import make_frame
    from oracle import geomae_oracle as O
    torch.set_num_threads(threads)
    kw = dict(geometry or {})
    if blocks:
        kw.update(enc_blocks=blocks[0], dec_blocks=blocks[1])
    cfg = O.PathConfig(**kw)
    frames = [make_frame(seed + s, **frame_kw) for s in range(samples)]
    params = {k: v.requires_grad_(True) for k, v in O.init_params(cfg, 0).items()}
    rows, _, _ = O.unique_rows(O.batch_voxelize(frames, cfg.voxel_size, cfg.pc_range))
    keep, mask = O.vanilla_mask_ids(rows, len(frames), cfg.mask_ratio, seed)
    t0 = time.perf_counter()
    losses, _, _ = O.forward_train(params, frames, cfg, keep, mask)
    sum(losses.values()).backward()
    return time.perf_counter() - t0, len(frames)


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference is Python on mmcv/mmdet/spconv/
    torch_scatter, none of which exist here, so the timed thing is the oracle port (kind "port"),
    pinned against the unmodified reference by tests/golden.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    samples = 1                      # bounded sample: one frame per step keeps K+W steps within minutes
    wl = workload_of(args)
    WORKLOAD = wl["label"]
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(samples, wl["frame"], threads, geometry=wl["geometry"])
    times, t0 = [], time.perf_counter()
    for i in range(max(1, args.steps)):
        t, n = cpu_reference_step(samples, wl["frame"], threads, seed=10 + i, geometry=wl["geometry"])
        times.append(t / n)
        if time.perf_counter() - t0 > 150.0:        # keep the whole arm within a few minutes
            break
    sec_per_frame = float(np.mean(times))
    v = 1.0 / sec_per_frame
    sample = f"{len(times)} steps x {samples} frame (fwd+bwd, full 6+2+2-block model, no optimiser), {threads} threads"
    print(json.dumps(dict(
        metric=METRIC, value=v, unit="frames/s", n_gpus=args.gpus, steps=len(times), warmup=min(args.warmup, 1),
        ms_per_step=1e3 * sec_per_frame * samples, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="f32", data="synthetic", impl="reference",
        config=dict(workload=WORKLOAD, samples_per_step=samples, sweeps=wl["frame"].get("sweeps", 1)),
        cpu_baseline=dict(value=v, unit="frames/s", cores=threads, kind="port", sample=sample),
        e2e=dict(value=v, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("GEOMAE_ALLOC_CONF"):          # experiment switch for the caching allocator
        torch.cuda.memory._set_allocator_settings(os.environ["GEOMAE_ALLOC_CONF"])
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))   # a hang must not eat the box
    torch.backends.cuda.matmul.allow_tf32 = False        # the reference trains fp32 with TF32 off (tools/train.py:24-25)
    torch.backends.cudnn.allow_tf32 = False

    import geomae_b200  # noqa: F401
    from geomae_b200 import lib as L
    from geomae_b200.registry import Config, build_model
    from geomae_b200.train import FlatTrainer

    from geomae_b200.workloads import with_geometry
    wl = workload_of(args)
    WORKLOAD = wl["label"]
    cfg = Config.fromfile(OWN_CFG)
    torch.manual_seed(0)
    model = build_model(with_geometry(cfg.model, wl["geometry"])).to(dev).train()
    model.set_impl(args.sra_impl)
    opt = cfg.optimizer
    trainer = FlatTrainer(model, lr=opt["lr"], betas=opt["betas"], weight_decay=opt["weight_decay"],
                          max_grad_norm=cfg.optimizer_config["grad_clip"]["max_norm"])
    K, W, S = args.steps, args.warmup, wl["samples_per_gpu"]
    pool = make_batches(rank, min(K + W, 8), S, wl["frame"])
    host = [[torch.from_numpy(f).pin_memory() for f in b] for b in pool]
    resident = [[f.to(dev) for f in b] for b in host]
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed_loop(step_fn, n, tag=None):
        """ONE pair of events around the n steps (the L2 flushes sit inside the timed region: the input stage of step
        i+1 overlaps the tail of step i, so per-step brackets would no longer add up to the job's time)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.collect()
        gc.disable()            # no cyclic-GC pause in the middle of a step (collected between loops instead)
        t0 = time.perf_counter()
        e0.record()
        for i in range(n):
            flush.zero_()                                   # evict L2 between steps
            step_fn(i)
        e1.record()
        if tag:
            host_ms[tag] = (time.perf_counter() - t0) * 1e3 / n      # how long the host needed to enqueue a step
        torch.cuda.synchronize()
        gc.enable()
        return e0.elapsed_time(e1)

    last = {}
    model.keep_targets = True

    def attention_pair_counts():
        """sum over windows of L^2 for the encoder / decoder token sets of the last step (both shifts)."""
        from geomae_b200.windows import WindowLayout
        tg = model.last_targets
        out = {}
        for name, rows in (("enc", tg["ids_keep"]), ("dec", torch.cat([tg["ids_keep"], tg["ids_mask"]]))):
            lay = WindowLayout.from_pillars(model.backbone.spec, tg["pillar_batch"], rows)
            for s in range(lay.spec.n_shifts):
                nw = int(lay.n_windows[s])
                ln = torch.diff(lay.win_ptr[s, :nw + 1]).double()
                out[name, s] = float((ln * ln).sum())
        return out

    def step_resident(i):
        last["loss"] = trainer.train_step(resident[i % len(resident)])[0]

    loss_pinned = torch.empty(2, dtype=torch.float32).pin_memory()
    loss_event = [None, None]

    def read_back(i, loss, final):
        """Device->host read of every step's loss inside the timed region, without draining the device each step: the
        value goes to pinned memory asynchronously and the host consumes it one step later (the last one before the
        closing synchronise), as a training loop's logger does."""
        slot = i & 1
        loss_pinned[slot:slot + 1].copy_(loss.reshape(1), non_blocking=True)
        loss_event[slot] = torch.cuda.Event()
        loss_event[slot].record()
        for s_ in ((slot ^ 1, slot) if final else (slot ^ 1,)):
            if loss_event[s_] is not None:
                loss_event[s_].synchronize()
                last["loss_host"] = float(loss_pinned[s_])
                loss_event[s_] = None

    def step_host(i):
        last["loss"] = trainer.train_step_from_host(host[i % len(host)])[0]
        read_back(i, last["loss"], i == K - 1)

    aug_rng = np.random.RandomState(1000 + rank)

    def step_host_augmented(i):
        """step_host with the train pipeline's GlobalRotScaleTrans / RandomFlip3D / PointsRangeFilter done on the
        device in front of the scatter: fresh draws every step, as the reference's workers do."""
        from geomae_b200.data import draw_augmentation
        batch = host[i % len(host)]
        augs = [draw_augmentation(aug_rng) for _ in batch]
        last["loss"] = trainer.train_step_from_host(batch, augs=augs)[0]
        read_back(i, last["loss"], i == K - 1)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                           # before the warm-up: nvidia-smi's start-up (NVML init) stalls launches
        sampler.wait_ready()
    for i in range(3 * len(resident)):            # allocator priming: every distinct synthetic batch, free-running (several
        step_resident(i)                          # steps in flight), so the timed steps never call cudaMalloc (untimed setup)
    for i in range(W):
        step_resident(i)
    barrier()
    import ctypes as C
    if not args.list_only:
        # K steps with CUDA events around every C-ABI call / stack kernel (roofline + kernel table).  The events cost
        # ~5-8 %, so this pass is separate from the headline loop; it runs FIRST and doubles as extra warm-up (the first
        # timed loop of a fresh process measured up to 15 % slow on this host-launch-bound step).
        L.start_timing()
        L.run("profile_enable", 1)
        Kp = min(K, 64)                                # the C side keeps 8192 spans (~106 per step)
        ms_prof = timed_loop(step_resident, Kp)
        per_call = L.stop_timing()
        prof_ms, prof_n, prof_fl, prof_by = (C.c_double * 5)(), (C.c_int64 * 5)(), (C.c_double * 5)(), (C.c_double * 5)()
        L.run("profile_read", prof_ms, prof_n, prof_fl, prof_by)
        L.run("profile_enable", 0)
        prof_ms, prof_n, prof_fl, prof_by = list(prof_ms), list(prof_n), list(prof_fl), list(prof_by)
        barrier()
    for i in range(max(W, 3)):                    # settle again after the instrumented pass (its 30 k events were just freed)
        step_resident(i)
    barrier()
    L.reset_call_counts()
    sampler.mark()
    ms = timed_loop(step_resident, K, "resident")  # the headline number: no per-kernel instrumentation
    sampler.mark()
    launches = L.launch_count()
    barrier()
    if args.list_only:
        if rank == 0:
            sampler.stop()
            print(json.dumps(dict(list_only=True, ms_per_step=ms / K, gpu_launches=launches)))
        if world > 1:
            dist.destroy_process_group()
        return
    for i in range(max(W, len(host))):            # warm-up + allocator priming of the host-fed path (input-stream pool)
        step_host(i)
    barrier()
    sampler.mark()
    ms_e2e = timed_loop(step_host, K, "e2e")
    sampler.mark()
    barrier()
    # bf16-mode vs parity-mode (bf16x3 + fp32 attention) loss on one identical batch, same weights, same mask split
    loss_delta = None
    if args.sra_impl != "tc3":      # every rank runs it: the VFE's synchronised BatchNorm is a collective
        with torch.no_grad():
            tg = model.last_targets
            ids = (tg["ids_keep"], tg["ids_mask"])
            batch = resident[(K - 1) % len(resident)]
            out = {}
            for impl in (args.sra_impl, "tc3"):
                model.set_impl(impl)
                model.forward_train(points=batch, img_metas=None, ids=ids)
                out[impl] = model.last_loss_vector.double().sum().item() if getattr(model, "last_loss_vector", None) is not None else None
            model.set_impl(args.sra_impl)
        if out[args.sra_impl] is not None and out["tc3"]:
            loss_delta = dict(loss=out[args.sra_impl], loss_tc3=out["tc3"],
                              rel=abs(out[args.sra_impl] - out["tc3"]) / abs(out["tc3"]))
    for i in range(len(host)):
        step_host_augmented(i)
    # ^ warm-up; placed after the loss-delta block: that one re-uses the last un-augmented step's mask split
    barrier()
    ms_e2e_aug = timed_loop(step_host_augmented, K, "e2e_aug")
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, ms_e2e, ms_e2e_aug], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e, ms_e2e_aug = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    frames = K * S * world
    h2d = sum(f.numel() * 4 for f in host[0])
    # roofline of the dominant hand-written kernel family, timed live with CUDA events inside the stack executor
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    tens_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured" if peaks else "fallback"
    fused = args.sra_impl == "tc1"
    attn = ("k_sra_tc_fwd", "k_sra_tc_bwd") if fused else ("k_sra_fwd", "k_sra_bwd_q+k_sra_bwd_kv")
    fam_names = ["k_sra_chain_fwd+k_sra_chain_bwd" if fused else "k_tc_linear", "k_wgrad_tma" if fused else "k_tc_wgrad",
                 attn[0], attn[1], "k_ln_bwd"]
    roof, fam_table = None, {}
    # DRAM traffic per launch of the SRA kernels from the committed ncu --set full capture of this command
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == args.workload and tj.get("sra_impl") == args.sra_impl:
            traffic = tj.get("kernels", {})
    except (OSError, ValueError):
        pass
    if prof_ms is not None and sum(prof_ms) > 0:
        sq = attention_pair_counts()
        bb = model.backbone
        n_enc, n_dec = len(bb.encoder_blocks), len(bb.decoder_centroid_blocks) + len(bb.decoder_density_blocks)
        pairs = (n_enc * (sq["enc", 0] + sq["enc", 1]) + n_dec * (sq["dec", 0] + sq["dec", 1])) * Kp   # all layers, Kp steps
        prof_fl[2] = pairs * bb.nhead[0] * 64.0     # fwd: QK^T + PV = 2 products x 16 x 2 flop per (i,j,head)
        prof_fl[3] = pairs * bb.nhead[0] * 160.0    # bwd: 5 products
        for f, nm in enumerate(fam_names):
            if prof_n[f]:
                sec = prof_ms[f] * 1e-3
                fam_table[nm] = dict(ms_per_step=prof_ms[f] / Kp, launches_per_step=prof_n[f] / Kp,
                                     avg_launch_us=1e3 * prof_ms[f] / prof_n[f],
                                     algorithmic_gflop_per_step=prof_fl[f] / Kp / 1e9,
                                     tflops=prof_fl[f] / sec / 1e12 if prof_fl[f] else None,
                                     tensor_frac=prof_fl[f] / sec / 1e12 / tens_peak if prof_fl[f] else None,
                                     algorithmic_gb_per_step=prof_by[f] / Kp / 1e9, gbs=prof_by[f] / sec / 1e9,
                                     hbm_frac=prof_by[f] / sec / 1e9 / hbm_peak)
        # the SRA layer (projections + attention + FFN) is bounded by the TENSOR pipe:
        # achieved = algorithmic flops of the family's launches / their CUDA-event time, peak = the measured sustained
        # bf16 cuBLAS rate.  The byte view of the same launches is kept as a secondary key (hbm_view).
        top = max(range(4), key=lambda f: prof_ms[f])
        ach = prof_fl[top] / (prof_ms[top] * 1e-3) / 1e12
        tr = traffic.get(fam_names[top], {})
        roof = dict(kernel=fam_names[top], bound="tensor", achieved=ach, peak=tens_peak, unit="TFLOP/s", frac=ach / tens_peak,
                    traffic=tr.get("dram_bytes_per_launch"), traffic_source=tr.get("source"),
                    peak_source=peak_src + " (bf16_tflops_sustained: the kernel is timed inside a long step)",
                    avg_launch_ms=prof_ms[top] / prof_n[top], launches=int(prof_n[top]),
                    algorithmic_flops_per_launch=prof_fl[top] / prof_n[top],
                    hbm_view=dict(algorithmic_bytes_per_launch=prof_by[top] / prof_n[top],
                                  achieved_gbs=prof_by[top] / (prof_ms[top] * 1e-3) / 1e9, peak_gbs=hbm_peak,
                                  frac=prof_by[top] / (prof_ms[top] * 1e-3) / 1e9 / hbm_peak),
                    whole_step=dict(algorithmic_gflop=sum(prof_fl[:4]) / Kp / 1e9,
                                    tflops=sum(prof_fl[:4]) / Kp / (ms / K * 1e-3) / 1e12,
                                    tensor_frac=sum(prof_fl[:4]) / Kp / (ms / K * 1e-3) / 1e12 / tens_peak),
                    note="timed with CUDA events on the launching stream around every launch of the family in a separate "
                         "instrumented pass over the same K steps; launches of concurrent streams overlap, so family "
                         "shares can add up to more than 1; the ncu launch list of the same command is under profiles/",
                    share_of_step=prof_ms[top] / ms_prof, instrumented_ms_per_step=ms_prof / Kp)
    aux = hbm_kernel_rooflines(model, pool, dev, hbm_peak, peak_src)
    line = dict(
        metric=METRIC, value=frames / (ms * 1e-3), unit="frames/s", n_gpus=world, steps=K, warmup=W,
        ms_per_step=ms / K, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype={"tc1": "bf16", "tc3": "bf16x3 (fp32-equivalent)", "glue": "f32"}[args.sra_impl],
        data="synthetic",
        config=dict(workload=WORKLOAD, workload_key=args.workload, samples_per_gpu=S, sweeps=wl["frame"].get("sweeps", 1),
                    points_per_frame=int(pool[0][0].shape[0]),
                    parallelism=f"dp{world}", l2="flushed between steps (192 MiB write on the compute stream, inside the timed region)",
                    sra_impl=args.sra_impl,
                    precision={"tc1": "SRA GEMMs bf16 operands / fp32 TMEM accumulate; activations, LayerNorm, softmax, VFE, targets, losses fp32",
                               "tc3": "SRA GEMMs bf16x3 split on tensor cores (fp32-equivalent, loss parity <=1e-4); rest fp32",
                               "glue": "fp32 library GEMMs, TF32 off"}[args.sra_impl]),
        e2e=dict(value=frames / (ms_e2e * 1e-3), unit="frames/s", h2d_bytes_per_step=h2d,
                 d2h_bytes_per_step=4 + 4 * (5 + S), ms_per_step=ms_e2e / K,
                 reads="every step: the scatter's pillar totals (blocking, on the input stream) and the loss (copied to "
                       "pinned memory asynchronously, consumed by the host one step later, the last before the closing "
                       "synchronise)"),
        # the same end-to-end step with the train pipeline's rotate / scale / flip / range filter done on the device
        # in front of the scatter (geomae_augment_filter; fresh draws per step, one extra 20-byte offsets read)
        e2e_with_device_augmentation=dict(value=frames / (ms_e2e_aug * 1e-3), unit="frames/s",
                                          ms_per_step=ms_e2e_aug / K, h2d_bytes_per_step=h2d,
                                          d2h_bytes_per_step=4 + 4 * (5 + S) + 4 * (S + 1)),
        host_enqueue_ms_per_step=host_ms, gpu_launches=launches, clocks=clocks, roofline=roof, kernel_families=fam_table, hbm_kernels=aux,
        kernel_ms_per_step={k: round(v[0] / Kp, 4) for k, v in sorted(per_call.items(), key=lambda kv: -kv[1][0])},
        loss=last.get("loss_host"), loss_delta_vs_tc3=loss_delta)
    if (args.workload == "dense" or args.length_bins) and world == 1:
        line["sra_length_bins"] = sra_length_bin_sweep(dev, tens_peak)
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1

        def sample_cpu(budget, **kw):
            cpu_reference_step(1, wl["frame"], threads, geometry=wl["geometry"], **kw)      # warm-up (thread pools, allocator)
            t_cpu, n_cpu, t0 = 0.0, 0, time.perf_counter()
            while time.perf_counter() - t0 < budget and n_cpu < 64:       # bounded sample of host work
                sec, n = cpu_reference_step(1, wl["frame"], threads, seed=2 + n_cpu, geometry=wl["geometry"], **kw)
                t_cpu += sec
                n_cpu += n
            return n_cpu, t_cpu
        n_cpu, t_cpu = sample_cpu(12.0)
        line["cpu_baseline"] = dict(value=n_cpu / t_cpu, unit="frames/s", cores=threads, kind="port",
                                    sample=f"{n_cpu} single-frame steps (fwd+bwd of the full 6+2+2-block model, no optimiser) "
                                           f"of the oracle port in {t_cpu:.1f} s, torch CPU fp32 on {threads} threads")
        if args.workload == "nus":      # BASELINE.json configs[0]: single frame, 1 SST block, CPU reference path
            n0, t0_ = sample_cpu(6.0, blocks=(1, 1))
            line["cpu_baseline"]["config0"] = dict(value=n0 / t0_, unit="frames/s", cores=threads, kind="port",
                                                   sample=f"{n0} single-frame steps of the 1-encoder-block / 1-decoder-block "
                                                          f"model (BASELINE.json configs[0]) in {t0_:.1f} s")
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
