"""Data-parallel training step for the emulator models (one process per GPU).

Replaces what Lightning's Trainer + DDP + torch.optim.Adam do around the reference's
`training_step` (main_final.py:556-561, :737-747): forward -> MSE -> backward ->
gradient all-reduce (mean over ranks, NCCL over NVLink/NVSwitch; skipped at world size 1) -> Adam.
The whole step is captured once into a CUDA graph and replayed (the reference's step is ~8 000
ATen launches; ours is a few hundred kernels whose launch cost the graph removes).

Samples are independent (GroupNorm is per sample), so the path shards by batch with ONE exchange:
the all-reduce of the flat gradient buffer (`post_conv`'s never-used parameters sit at its tail
and are excluded — SURVEY F5)."""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist

from . import ops
from ._lib import lib
from .optim import FusedAdam
from .parallel import allreduce_flat_grads


class TrainStep:
    def __init__(self, model: torch.nn.Module, x_shape, y_shape, lr: float = 5e-4, weight_decay: float = 0.0,
                 use_graph: bool = True, process_group=None, device: Optional[torch.device] = None):
        self.model = model
        self.device = device or next(model.parameters()).device
        unused = list(model.post_conv.parameters()) if hasattr(model, "post_conv") else []
        self.opt = FusedAdam(model.parameters(), lr=lr, weight_decay=weight_decay, unused=unused)
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.pg = process_group
        self.x = torch.zeros(x_shape, device=self.device, dtype=torch.float32)
        self.y = torch.zeros(y_shape, device=self.device, dtype=torch.float32)
        self.loss = torch.zeros((), device=self.device, dtype=torch.float32)
        self.use_graph = use_graph
        self.plan = ops.PackPlan()          # resident packed weights, refreshed by one launch per step
        fg = self.opt.flat_grad
        self.plan.grad_range = (fg.data_ptr(), fg.data_ptr() + fg.numel() * 4)
        # weight-gradient kernels on a second stream, forked / joined inside the step (also under graph capture): nothing
        # in the backward chain depends on them, and since the per-image tails stopped filling the SMs' shared memory
        # for their whole duration the branches do overlap (2.07 -> 1.99 ms per step on B200).  PCM_SIDE_STREAM=0
        # keeps everything on one stream.
        self.side = torch.cuda.Stream(device=self.device) if os.environ.get("PCM_SIDE_STREAM", "1") != "0" else None
        # Gradient buckets (world > 1), in the order backward completes them.  A bucket is a range of the flat gradient
        # buffer plus the name of the backward hook at which all of its gradients have been enqueued:
        #   [convlstm .. head]  at "encoder_boundary" (backward reaches the ConvLSTM input)
        #   [enc2 .. enc4]      at "enc1_boundary"    (backward reaches enc2's input)
        #   [enc1]              at the end of backward (14 KB: the only all-reduce that stays exposed)
        # At each hook the bucket's packed gradients are folded, and on a third stream its all-reduce is issued and
        # followed by the Adam update of exactly that parameter range — both overlap the rest of backward.
        # PCM_OVERLAP_ALLREDUCE=0: one all-reduce + one Adam launch after backward; PCM_GRAD_BUCKETS=2: the two-bucket
        # scheme of round 1 (no hook at the enc1 boundary).
        self.buckets = []            # [(hook name, lo, hi)]
        self.tail_bucket = None      # (lo, hi) reduced after backward
        self.split = 0
        # A single GPU uses the same buckets without the all-reduce (PCM_BUCKETS_SINGLE=0: one fold + one Adam launch at
        # the end): the folds and Adam updates of the two big buckets then run on the communication stream while
        # backward continues, and only enc1's fold and update stay behind the last weight gradient.
        single = self.world == 1 and os.environ.get("PCM_BUCKETS_SINGLE", "1") != "0"
        if ((self.world > 1 or single) and hasattr(model, "convlstm")
                and os.environ.get("PCM_OVERLAP_ALLREDUCE", "1") != "0"):
            off = lambda p: (p.main_grad.data_ptr() - fg.data_ptr()) // 4
            s_lstm = off(next(model.convlstm.parameters()))
            self.split = s_lstm
            self.buckets.append(("encoder_boundary", s_lstm, self.opt.n_reduced))
            lo = s_lstm
            if hasattr(model, "enc2") and os.environ.get("PCM_GRAD_BUCKETS", "3") != "2":
                s_enc2 = off(next(model.enc2.parameters()))
                if 0 < s_enc2 < s_lstm:
                    self.buckets.append(("enc1_boundary", s_enc2, s_lstm))
                    lo = s_enc2
            self.tail_bucket = (0, lo)
        self.comm_stream = torch.cuda.Stream(device=self.device) if self.buckets else None
        # models without any dropout site (the UNet family) skip the per-step advance of the device-side mask epoch
        self.has_dropout = (float(getattr(model, "p_drop", 0.0) or 0.0) > 0.0 or any(
            isinstance(m, (torch.nn.Dropout, torch.nn.Dropout2d, torch.nn.Dropout3d)) and m.p > 0.0 or
            isinstance(m, torch.nn.MultiheadAttention) and m.dropout > 0.0 for m in model.modules()))
        self._copy_stream = None
        self.graph = None
        self.launches_per_step = 0
        self.check_every = int(os.environ.get("PCM_CHECK_EVERY", "256"))    # poll the device error counter (0: never)
        self._steps_run = 0
        self._ragged_plans = {}
        self._windows = None                # data.WindowLoader while the window-index step is being built / run
        self.idx = None
        self.graph_windows = None

    # -- one eager step on the static buffers ---------------------------------------------------
    def _bucket_ready(self, lo: int, hi: int):
        """Backward has passed this bucket's hook: fold its packed gradients, then — on the communication stream, so that
        the rest of backward keeps running — all-reduce the range and apply Adam to it."""
        fg = self.opt.flat_grad
        # the communication stream waits for backward so far and for the weight gradients (side stream), then folds,
        # reduces and updates; the main stream does not wait for anybody (PCM_BUCKET_FOLD_MAIN=1: round-2 behaviour, the
        # fold on the main stream after joining the side stream)
        if os.environ.get("PCM_BUCKET_FOLD_MAIN", "0") == "1" and self.world > 1:
            ops.join_side()
            self.plan.unpack_grads(fg.data_ptr() + 4 * lo, fg.data_ptr() + 4 * hi)
            self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(fg[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
                self.opt.step_range(lo, hi, grad_scale=1.0 / self.world)
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream(self.device))
        if self.side is not None and ops.side_forked():
            self.comm_stream.wait_stream(self.side)
        with torch.cuda.stream(self.comm_stream):
            self.plan.unpack_grads(fg.data_ptr() + 4 * lo, fg.data_ptr() + 4 * hi)
            if self.world > 1:
                dist.all_reduce(fg[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
            self.opt.step_range(lo, hi, grad_scale=1.0 / self.world)

    def _forward_loss(self):
        if self._windows is not None:
            # device-resident record: the batch is B window indices; gather + zero left-pad +
            # normalisation + seasonal channels + NHWC staging are one kernel
            from .config import compute_dtype
            xs, y = self._windows.stage(self.idx, compute_dtype())
            if hasattr(self.model, "forward_loss"):
                return self.model.forward_staged(xs, self.idx.numel(), self._windows.seq_len, target=y)
            out = self.model.forward_staged(xs, self.idx.numel(), self._windows.seq_len)
            return ops.mse_loss(out, y)
        if hasattr(self.model, "forward_loss"):
            return self.model.forward_loss(self.x, self.y)          # head + loss fused (ops.HeadMSEFn)
        return ops.mse_loss(self.model(self.x), self.y)

    def _step_impl(self):
        # fresh dropout masks on every step, also under graph replay (the scalar seeds are frozen in the graph; the
        # kernels mix this device-side counter into them)
        if self.has_dropout:
            lib().call("pcm_dropout_epoch_advance", torch.cuda.current_stream().cuda_stream)
        self.opt.zero_grad()
        fg = self.opt.flat_grad
        overlap = bool(self.buckets)
        if overlap:
            self.opt.tick()                                   # one step-counter advance; Adam then runs per bucket
            for name, lo, hi in self.buckets:
                ops._GRAD_HOOKS[name] = (lambda lo=lo, hi=hi: self._bucket_ready(lo, hi))
        try:
            with ops.use_pack_plan(self.plan, self.side):
                # the one-launch weight re-pack runs beside the input staging (second stream); the first layer that takes
                # a packed weight joins it (ops.pack_weight joins on the first hit)
                with ops.side_stream():
                    self.plan.repack()
                self.plan.pending_join = ops.side_forked()
                loss = self._forward_loss()
                self.plan.pending_join = False
                loss.backward()
                ops.join_side()
                if overlap:
                    lo, hi = self.tail_bucket
                    self.plan.unpack_grads(fg.data_ptr() + 4 * lo, fg.data_ptr() + 4 * hi)
                else:
                    self.plan.unpack_grads()
        finally:
            for name, _, _ in self.buckets:
                ops._GRAD_HOOKS.pop(name, None)
        if overlap:
            lo, hi = self.tail_bucket
            if self.world > 1:
                dist.all_reduce(fg[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
            self.opt.step_range(lo, hi, grad_scale=1.0 / self.world)
            torch.cuda.current_stream(self.device).wait_stream(self.comm_stream)     # the other buckets' updates are done
        else:
            scale = allreduce_flat_grads(fg, self.opt.n_reduced, self.pg)
            self.opt.step(grad_scale=scale)
        self.loss.copy_(loss.detach())

    def _snapshot(self):
        bufs = [b for b in self.model.buffers()]
        return (self.opt.flat_param.clone(), self.opt.exp_avg.clone(), self.opt.exp_avg_sq.clone(),
                self.opt.state.clone(), [b.clone() for b in bufs], bufs)

    def _restore(self, snap):
        fp, m, v, st, saved, bufs = snap
        self.opt.flat_param.copy_(fp); self.opt.exp_avg.copy_(m); self.opt.exp_avg_sq.copy_(v); self.opt.state.copy_(st)
        for b, sb in zip(bufs, saved):
            b.copy_(sb)

    def warmup_and_capture(self, warmup: int = 3, restore: bool = True):
        """Eager warm-up on a side stream (also sizes the allocator), then capture the graph.
        Warm-up steps run real optimisation steps on whatever is in the static buffers; with `restore` (default) the
        parameters, the Adam state and the module buffers (BatchNorm running statistics) are snapshotted before and put
        back afterwards, so training starts from the weights the caller loaded, at step 0."""
        snap = self._snapshot() if restore else None
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for i in range(warmup):
                n0 = lib().launches
                self._step_impl()
                self.launches_per_step = lib().launches - n0
        torch.cuda.current_stream(self.device).wait_stream(s)
        torch.cuda.synchronize(self.device)
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            # The critical chain is captured on a HIGH-priority stream, the side / communication streams have the default
            # (lowest) priority: kernel nodes inherit it, so when a main-chain kernel and a weight-gradient kernel both
            # have CTAs pending the scheduler serves the chain first.
            with torch.cuda.graph(self.graph, stream=self._capture_stream()):
                self._step_impl()
            torch.cuda.synchronize(self.device)
        if snap is not None:
            self._restore(snap)
        self.check_kernels()

    def _capture_stream(self):
        """High-priority stream for the captured critical chain (kernel nodes inherit the priority; the side and
        communication streams keep the default = lowest one).  Measured on B200: 1.78 -> 1.64 ms per step.
        PCM_MAIN_PRIORITY=0: torch's default capture stream."""
        prio = int(os.environ.get("PCM_MAIN_PRIORITY", "-1"))
        return torch.cuda.Stream(device=self.device, priority=prio) if prio != 0 else None

    def check_kernels(self):
        """The tensor-core kernels report a pipeline time-out (a bounded mbarrier wait that expired) only through a
        device counter and skip their epilogue; poll it (one 4-byte D2H, synchronising) and refuse to go on with
        silently wrong activations.  Called after capture and every `check_every` steps from `run`."""
        n = int(lib()._fn["pcm_tc_error_count"]())
        if n != 0:
            raise RuntimeError(f"pcm_b200: {n} tensor-core pipeline time-out(s) reported by the device "
                               "(pcm_tc_error_count) — results of the affected launches are invalid")

    def capture_windows(self, loader, warmup: int = 2, restore: bool = True):
        """Second captured step whose input is `self.idx` — B window indices into `loader`'s HBM-resident record
        (data.WindowLoader) — instead of a staged (B, T, C, H, W) batch: per step only the indices cross PCIe."""
        B = self.x.shape[0]
        self.idx = torch.zeros(B, dtype=torch.int64, device=self.device) if self.idx is None else self.idx
        snap = self._snapshot() if restore else None
        plan, self.plan = self.plan, ops.PackPlan()
        self.plan.grad_range = plan.grad_range
        self._windows = loader
        try:
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    self._step_impl()
            torch.cuda.current_stream(self.device).wait_stream(s)
            torch.cuda.synchronize(self.device)
            if self.use_graph:
                self.graph_windows = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_windows, stream=self._capture_stream()):
                    self._step_impl()
                torch.cuda.synchronize(self.device)
            self._plan_windows = self.plan
        finally:
            self.plan = plan
            self._windows = None
        self._loader = loader
        if snap is not None:
            self._restore(snap)

    def step_windows(self, idx: torch.Tensor) -> torch.Tensor:
        """One optimisation step on the windows ending at `idx` (B,) int64 (host pinned or device)."""
        if tuple(idx.shape) != tuple(self.idx.shape):
            raise ValueError(f"step_windows: expected {tuple(self.idx.shape)} indices, got {tuple(idx.shape)}")
        self.idx.copy_(idx, non_blocking=True)
        if self.graph_windows is not None:
            self.graph_windows.replay()
        else:
            plan, self.plan, self._windows = self.plan, self._plan_windows, self._loader
            try:
                self._step_impl()
            finally:
                self.plan, self._windows = plan, None
        self._steps_run += 1
        return self.loss

    def reset_optimizer_state(self):
        self.opt.exp_avg.zero_()
        self.opt.exp_avg_sq.zero_()
        self.opt.state.zero_()

    def load_batch(self, x: torch.Tensor, y: torch.Tensor):
        """Copy a batch (host pinned or device) into the static input buffers (async).  The step is captured for ONE
        batch shape: a different shape (e.g. the last, smaller batch of an epoch — the reference's DataLoader does not
        drop it, main_final.py:486-492) must go through `step_ragged`, which pads it and masks the loss."""
        if tuple(x.shape) != tuple(self.x.shape) or tuple(y.shape) != tuple(self.y.shape):
            raise ValueError(f"TrainStep was built for x {tuple(self.x.shape)} / y {tuple(self.y.shape)}, got "
                             f"{tuple(x.shape)} / {tuple(y.shape)}; use step_ragged() for a smaller final batch or "
                             "drop_last=True")
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)

    def run(self):
        """One optimisation step on whatever is in the static buffers; returns the device loss."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_impl()
        self._steps_run += 1
        if self.check_every and self._steps_run % self.check_every == 0:
            self.check_kernels()
        return self.loss

    def step_ragged(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        """The final, smaller batch of an epoch (b < B rows): one EAGER step on exactly those rows (same kernels, no
        graph: its shape differs from the captured one).  MSE is the mean over the b rows, as nn.MSELoss gives the
        reference.  Single-process only (under data parallelism every rank must issue the same collectives; use
        drop_last or pad the sampler like DistributedSampler does)."""
        if self.world > 1:
            raise RuntimeError("step_ragged: ragged batches are not supported under data parallelism (pad the sampler)")
        b = x.shape[0]
        if b == self.x.shape[0]:
            return self.step(x, y)
        if b == 0 or b > self.x.shape[0] or tuple(x.shape[1:]) != tuple(self.x.shape[1:]):
            raise ValueError(f"step_ragged: bad batch shape {tuple(x.shape)} for a TrainStep of {tuple(self.x.shape)}")
        xs, ys = self.x, self.y
        try:
            self.x = x.to(self.device, torch.float32, non_blocking=True)
            self.y = y.to(self.device, torch.float32, non_blocking=True)
            plan = self.plan
            if b not in self._ragged_plans:                      # own resident buffers: the captured plan's tables
                self._ragged_plans[b] = ops.PackPlan()           # must not change after capture
                self._ragged_plans[b].grad_range = plan.grad_range
            self.plan = self._ragged_plans[b]
            try:
                self._step_impl()
            finally:
                self.plan = plan
        finally:
            self.x, self.y = xs, ys
        return self.loss

    def step(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.load_batch(x, y)
        return self.run()

    # -- input prefetch: the host->device copy of the NEXT batch overlaps this step --------------------------
    def _alt_buffers(self):
        """Second pair of static input buffers and — under graph replay — a second captured step that reads them (same
        memory pool: the two graphs never run concurrently).  step_prefetch alternates between the pairs, so the next
        batch lands by H2D exactly where the next step reads it and no device-to-device move is left on the step's path."""
        self._x_alt = torch.empty_like(self.x)
        self._y_alt = torch.empty_like(self.y)
        self._graph_alt = None
        if self.graph is not None:
            torch.cuda.synchronize(self.device)
            self.x, self._x_alt = self._x_alt, self.x
            self.y, self._y_alt = self._y_alt, self.y
            try:
                g = torch.cuda.CUDAGraph()
                # capture enqueues nothing: parameters / optimizer state are untouched
                with torch.cuda.graph(g, pool=self.graph.pool(), stream=self._capture_stream()):
                    self._step_impl()
                self._graph_alt = g
            finally:
                self.x, self._x_alt = self._x_alt, self.x
                self.y, self._y_alt = self._y_alt, self.y
            torch.cuda.synchronize(self.device)
        self._free_alt = torch.cuda.Event()       # the step that last read the (currently) alternate pair has finished
        self._free_alt.record()
        self._free_cur = torch.cuda.Event()

    def step_prefetch(self, x_next: torch.Tensor, y_next: torch.Tensor) -> torch.Tensor:
        """Run one step on the batch already resident in the static buffers while (x_next, y_next) — pinned host
        tensors — are copied on a second stream into the ALTERNATE pair of static buffers, which the next call then
        trains on (the pairs, and the two captured graphs that read them, swap roles after every call; `load_batch` /
        `run` / `step` always address the current pair).  This is what a DataLoader with pinned memory and non_blocking
        copies does for the reference's training loop (main_final.py:291), made explicit: every call still moves one
        batch across PCIe, but off the critical path, and nothing is moved twice."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._staged = torch.cuda.Event()
            self._alt_buffers()
        main = torch.cuda.current_stream(self.device)
        loss = self.run()                                       # the step is launched FIRST: the host-side set-up of the
        self._free_cur.record(main)                             # copies below then overlaps it instead of delaying it
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._free_alt)        # its previous contents have been consumed
            self._x_alt.copy_(x_next, non_blocking=True)
            self._y_alt.copy_(y_next, non_blocking=True)
            self._staged.record()
        main.wait_event(self._staged)                           # the next step (any stream order) sees the staged batch
        self.x, self._x_alt = self._x_alt, self.x
        self.y, self._y_alt = self._y_alt, self.y
        self.graph, self._graph_alt = self._graph_alt, self.graph
        self._free_cur, self._free_alt = self._free_alt, self._free_cur
        return loss
