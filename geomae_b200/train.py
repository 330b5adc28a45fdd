"""One data-parallel pre-training step: forward_train -> backward -> gradient all-reduce ->
grad-clip + AdamW.  Stands in for the mmcv runner pieces the reference uses around the hot path
(EpochBasedRunner.run_iter / OptimizerHook / MMDistributedDataParallel): one
process per GPU, all parameters and gradients live in ONE flat fp32 buffer each, so the DDP
exchange is a single NCCL all-reduce over NVLink and the optimiser is a single fused kernel."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import lib as L


class FlatTrainer:
    # parameters whose gradients are complete once the backward pass has left the decoders (they are all-reduced while
    # the encoder / VFE backward still runs); everything else goes with the second bucket after backward
    EARLY_KEYS = ("backbone.decoder_", "backbone.cls_pred_")

    def __init__(self, model, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05, max_grad_norm=10.0,
                 no_decay_keys=("norm",), overlap_input=True, allocator_rounding=True, peer_gradients=True):
        self.model = model
        if allocator_rounding and torch.cuda.is_available() and not os.environ.get("PYTORCH_CUDA_ALLOC_CONF"):
            # pillar / token counts differ from step to step, so nearly every buffer of a step has a size never seen
            # before; with exact sizes the caching allocator keeps splitting blocks and falls back to cudaMalloc inside
            # the loop (measured: host time per step jumping from 3.3 to 4-20 ms on fresh augmentations).  Rounding
            # request sizes to 1/16 steps of a power of two makes them repeat (<= 6 % more memory).  Process-wide
            # PyTorch setting; an explicit PYTORCH_CUDA_ALLOC_CONF wins.
            setter = getattr(torch._C, "_accelerator_setAllocatorSettings", None) or torch.cuda.memory._set_allocator_settings
            setter("roundup_power2_divisions:16")
        self.overlap_input = overlap_input      # run the input stage on its own stream (see input_stream())
        named = [(k, p) for k, p in model.named_parameters() if p.requires_grad]     # frozen parameters stay outside
        is_nd = lambda k: any(s in k for s in no_decay_keys)                        # noqa: E731
        is_early = lambda k: k.startswith(self.EARLY_KEYS)                           # noqa: E731
        groups = [[(k, p) for k, p in named if not is_nd(k) and is_early(k)],        # decay, early
                  [(k, p) for k, p in named if not is_nd(k) and not is_early(k)],    # decay, late
                  [(k, p) for k, p in named if is_nd(k) and not is_early(k)],        # no decay, late
                  [(k, p) for k, p in named if is_nd(k) and is_early(k)]]            # no decay, early
        self.order = [kp for g in groups for kp in g]
        align = 64      # floats: every tensor starts 256-byte aligned (the kernels use 128-bit accesses)

        def padded(p):
            return (p.numel() + align - 1) // align * align
        sizes = [sum(padded(p) for _, p in g) for g in groups]
        self.n_decay = sizes[0] + sizes[1]
        self.n = sum(sizes)
        # flat layout [decay early | decay late | no-decay late | no-decay early]: the decayed prefix the optimiser kernel
        # needs, the late bucket contiguous, the early bucket = the two ends
        self.early_ranges = [(0, sizes[0]), (self.n - sizes[3], self.n)]
        self.late_range = (sizes[0], self.n - sizes[3])
        self.n_params = sum(p.numel() for _, p in self.order)
        dev = named[0][1].device
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.flat_param = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.shared_grads = None
        if self.world > 1 and peer_gradients and dev.type == "cuda" and not os.environ.get("GEOMAE_NO_PEER_EXCHANGE"):
            # the gradient buffer of every rank mapped on every rank: the DDP exchange becomes one reduce-scatter +
            # all-gather kernel over NVLink peer memory per bucket (peer.SharedGradients) instead of NCCL all-reduces
            from .peer import SharedGradients
            try:
                self.shared_grads = SharedGradients(dev, self.n)
            except Exception as e:      # noqa: BLE001
                import warnings
                warnings.warn(f"geomae_b200: peer-memory gradient exchange unavailable ({e}); using NCCL all-reduce")
        self.flat_grad = self.shared_grads.flat_grad if self.shared_grads is not None else \
            torch.zeros(self.n, dtype=torch.float32, device=dev)
        off = 0
        for _, p in self.order:
            n = p.numel()
            self.flat_param[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + n].view_as(p)
            p.grad = self.flat_grad[off:off + n].view_as(p)
            off += padded(p)
        self.exp_avg = torch.zeros_like(self.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.flat_param)
        self.partials = torch.zeros(1024, dtype=torch.float64, device=dev)
        self.stats = torch.zeros(2, dtype=torch.float32, device=dev)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        self.step_count = 0

    def zero_grad(self):
        self.flat_grad.zero_()

    def check_bindings(self):
        """The fused backward passes accumulate into the flat gradient buffer through captured pointers: a
        ``model.to()`` / ``.float()`` / ``zero_grad(set_to_none=True)`` after construction would silently detach them."""
        off, align = 0, 64
        base_p, base_g = self.flat_param.data_ptr(), self.flat_grad.data_ptr()
        for k, p in self.order:
            if p.data_ptr() != base_p + 4 * off or p.grad is None or p.grad.data_ptr() != base_g + 4 * off:
                raise RuntimeError(f"FlatTrainer: parameter {k} no longer lives in the flat buffers (re-create the trainer "
                                   "after moving / casting the model; never set .grad to None)")
            off += (p.numel() + align - 1) // align * align

    # ---- checkpoint / resume (mmcv CheckpointHook saves the optimizer every epoch, default_runtime.py:1,16-17)
    def state_dict(self):
        return dict(step_count=self.step_count, order=[k for k, _ in self.order], n=self.n, n_decay=self.n_decay,
                    exp_avg=self.exp_avg.detach().cpu().clone(), exp_avg_sq=self.exp_avg_sq.detach().cpu().clone(),
                    lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.weight_decay,
                    max_grad_norm=self.max_grad_norm)

    def load_state_dict(self, sd):
        if sd["order"] != [k for k, _ in self.order] or sd["n"] != self.n:
            raise RuntimeError("FlatTrainer.load_state_dict: parameter order / sizes differ from the checkpoint")
        self.step_count = int(sd["step_count"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps = sd["lr"], tuple(sd["betas"]), sd["eps"]
        self.weight_decay, self.max_grad_norm = sd["weight_decay"], sd["max_grad_norm"]

    def optimizer_step(self, lr=None):
        self.step_count += 1
        L.run("adamw_step", 
            L.ptr(self.flat_param), L.ptr(self.flat_grad), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq), self.n,
            self.n_decay, L.ptr(self.partials), 1.0 / self.world, self.max_grad_norm, self.lr if lr is None else lr,
            self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count, L.ptr(self.stats),
            L.stream_ptr(self.flat_param.device))

    def input_stream(self):
        """The side stream of the input stage (H2D copies, device augmentation, voxel scatter and the read of its
        totals).  Work queued here never waits for the compute stream, so the one host read a step needs (pillar
        counts size every later buffer) blocks for the input stage only and the host keeps enqueueing ahead."""
        st = self.__dict__.get("_input_stream")
        if st is None:
            # default priority on purpose: a high-priority input stream measured slightly better medians but rare
            # 50-60 ms host stalls inside its allocations (tools/step_jitter_probe.py); at equal priority there are none
            st = self.__dict__["_input_stream"] = torch.cuda.Stream(self.flat_param.device)
        return st

    def train_step(self, points, ids=None, lr=None, ready=None):
        """points: list of [N_i, C] CUDA tensors (one rank's samples).  Returns (total loss, loss dict).

        The scatter runs on ``input_stream()``: ``points`` must be complete when this is called (synchronised, produced
        on ``input_stream()``, or guarded by ``ready``, a torch.cuda.Event recorded after whatever produced them)."""
        if self.step_count == 0:
            self.check_bindings()
        if hasattr(self.model, "extract_feat"):
            side = self.input_stream() if self.overlap_input else None
            if ready is not None:
                (side or torch.cuda.current_stream(self.flat_param.device)).wait_event(ready)
            self.model.scatter_stream = side
        prev = L.pin_stream(self.flat_param.device)
        try:
            return self._train_step(points, ids, lr)
        finally:
            L.unpin_stream(prev)

    def _train_step(self, points, ids, lr):
        self.zero_grad()
        self.model.last_loss_vector = None
        losses = self.model.forward_train(points=points, img_metas=None, ids=ids)
        vec = getattr(self.model, "last_loss_vector", None)
        total = vec.sum() if vec is not None else sum(losses.values())     # one reduction instead of six adds
        pending = []
        sg = self.shared_grads
        if self.world > 1:
            # bucket 1 (decoders + heads) starts its exchange from inside backward, as soon as the gradient reaching the
            # encoder output exists, and runs on its own stream under the encoder / VFE backward
            def early(_grad):
                if sg is not None:
                    here = torch.cuda.current_stream(self.flat_grad.device)
                    sg.stream.wait_stream(here)
                    with torch.cuda.stream(sg.stream), L.stream_override(sg.stream):
                        sg.exchange(sg.early, self.early_ranges)
                    pending.append(None)
                    return
                for a, b in self.early_ranges:
                    if b > a:
                        pending.append(dist.all_reduce(self.flat_grad[a:b], async_op=True))
            self.model.backbone.encoder_output_hook = early
        total.backward()
        if self.world > 1:
            self.model.backbone.encoder_output_hook = None
            if not pending:                          # the hook never fired (model without the backbone hook point)
                early(None)
            if sg is not None:
                sg.exchange(sg.late, [self.late_range])
                torch.cuda.current_stream(self.flat_grad.device).wait_stream(sg.stream)
            else:
                a, b = self.late_range
                pending.append(dist.all_reduce(self.flat_grad[a:b], async_op=True))
                for w in pending:
                    w.wait()                         # stream-level wait: the optimiser kernel is ordered after NCCL
        self.optimizer_step(lr)
        return total.detach(), losses

    def train_step_from_host(self, host_points, ids=None, lr=None, augs=None, point_cloud_range=None):
        """host_points: list of pinned CPU tensors (one frame each) or ``data.RawSweeps`` (a sample still split into its
        key frame and raw earlier sweeps: merged on the device by ``geomae_sweep_merge``, which does what
        LoadPointsFromMultiSweeps does on the reference's loader workers); the H2D copies are issued on
        ``input_stream()``.

        ``augs`` (one ``data.Augmentation`` per frame, e.g. from ``data.draw_augmentation``) runs the train pipeline's
        GlobalRotScaleTrans / RandomFlip3D / PointsRangeFilter on the device first (``geomae_augment_filter``, one call
        for the batch; configs/mae_sst/…6x_1e-5.py:180-190): the raw frames land in one buffer, are transformed,
        filtered and compacted there, and the per-frame survivor counts come back in one small read (the scatter stage
        sizes its buffers from the host-side point count).  Copies and the data step are queued on ``input_stream()``."""
        dev = self.flat_param.device
        main = torch.cuda.current_stream(dev)
        side = self.input_stream() if self.overlap_input else main
        from .data import RawSweeps, augment_filter, sweep_merge
        raw_sweeps = any(isinstance(p, RawSweeps) for p in host_points)
        with torch.cuda.stream(side):
            if augs is None and not raw_sweeps:
                pts = [p.to(dev, non_blocking=True) for p in host_points]
            else:
                # every segment (a whole frame, or one sweep file of a RawSweeps sample) lands in ONE device buffer
                segs, seg_par, first_seg = [], [], [0]
                for p in host_points:
                    if isinstance(p, RawSweeps):
                        segs += [a if torch.is_tensor(a) else torch.from_numpy(a) for a in p.arrays]
                        seg_par.append(torch.as_tensor(p.params))
                    else:
                        segs.append(p)
                        seg_par.append(None)
                    first_seg.append(len(segs))
                raw = torch.empty((sum(t.shape[0] for t in segs), segs[0].shape[1]), dtype=torch.float32, device=dev)
                offs, at = [0], 0
                for t in segs:
                    raw[at:at + t.shape[0]].copy_(t, non_blocking=True)
                    at += t.shape[0]
                    offs.append(at)
                if raw_sweeps:
                    ident = torch.zeros(1, 16, dtype=torch.float64)
                    ident[0, 0] = ident[0, 4] = ident[0, 8] = 1.0
                    ident[0, 13] = -1.0
                    par = torch.cat([ident if q is None else q for q in seg_par]).to(dev, non_blocking=True)
                    seg_off = torch.tensor(offs, dtype=torch.int32).to(dev, non_blocking=True)
                    raw, seg_off = sweep_merge(raw, seg_off, par)
                    frame_off = seg_off[torch.tensor(first_seg, dtype=torch.int64)]      # a sample starts at its first segment
                else:
                    frame_off = torch.tensor(offs, dtype=torch.int32).to(dev, non_blocking=True)
                if augs is not None:
                    rng = point_cloud_range if point_cloud_range is not None else self.model.point_cloud_range
                    raw, frame_off = augment_filter(raw, frame_off.contiguous(), augs, rng)
                o = frame_off.tolist()     # waits for the copies and the data-step kernels on the input stream only
                pts = [raw[o[b]:o[b + 1]] for b in range(len(host_points))]
        if side is not main and not hasattr(self.model, "extract_feat"):
            main.wait_stream(side)
        for t in pts:
            t.record_stream(main)
        return self.train_step(pts, ids=ids, lr=lr)


def cyclic_lr(base_lr, it, max_iters, target_ratio=(100, 1e-3), cyclic_times=1, step_ratio_up=0.1):
    """mmcv CyclicLrUpdaterHook as configured by configs/_base_/schedules/cosine_2x.py:10-15 (by_epoch=False, cosine
    annealing): lr rises from base_lr to base_lr*target_ratio[0] over the first step_ratio_up of a cycle, then falls to
    base_lr*target_ratio[1]; ``it`` counts iterations from 0."""
    import math
    period = max_iters // cyclic_times
    up = int(step_ratio_up * period)
    cur = it % period
    if cur < up:
        start, end, frac = 1.0, target_ratio[0], cur / up
    else:
        start, end, frac = target_ratio[0], target_ratio[1], (cur - up) / (period - up)
    cos_out = math.cos(math.pi * frac) + 1.0
    return base_lr * (end + 0.5 * (start - end) * cos_out)
