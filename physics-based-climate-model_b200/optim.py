"""Flat-buffer fused Adam (torch.optim.Adam semantics as configured at main_final.py:737-747).

All parameters are re-homed as views of one contiguous fp32 buffer, their gradients (`main_grad`,
also exposed as `.grad`) as views of a second one, so that
  * the backward kernels accumulate weight gradients straight into the flat buffer,
  * the data-parallel gradient all-reduce is ONE NCCL call over [0, n_reduced),
  * the optimizer is ONE kernel launch, and the step counter lives on the device so a captured
    CUDA graph advances it on every replay.
Parameters listed in `unused` (e.g. AttUNetConvLSTM.post_conv) are placed at the tail
of the buffers: they are excluded from the all-reduce and keep zero gradients."""
from __future__ import annotations

from typing import Iterable, List

import torch

from ._lib import lib


class FusedAdam:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr=5e-4, betas=(0.9, 0.999), eps=1e-8,
                 weight_decay=0.0, unused: Iterable[torch.nn.Parameter] = ()):
        unused_ids = {id(p) for p in unused}
        allp = [p for p in params if p.requires_grad]
        self._torch_order = list(allp)                  # torch.optim.Adam numbers parameters in this order
        used = [p for p in allp if id(p) not in unused_ids]
        tail = [p for p in allp if id(p) in unused_ids]
        self.params: List[torch.nn.Parameter] = used + tail
        if not self.params:
            raise ValueError("FusedAdam got no parameters")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdam needs CUDA parameters (no CPU fallback)")
        al = lambda n: (n + 3) // 4 * 4                  # 16-byte aligned slices
        self.n_reduced = sum(al(p.numel()) for p in used)
        total = self.n_reduced + sum(al(p.numel()) for p in tail)
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=dev, dtype=torch.float32)
        self.state = torch.zeros(4, device=dev, dtype=torch.float32)     # [step, 1-b1^t, 1-b2^t, -]
        off = 0
        self._offsets = {}
        for p in self.params:
            n = p.numel()
            self._offsets[id(p)] = off
            self.flat_param[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_param[off:off + n].view(p.shape)
            p.main_grad = self.flat_grad[off:off + n].view(p.shape)
            p.grad = p.main_grad
            off += al(n)
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay

    def zero_grad(self):
        self.flat_grad.zero_()

    def step(self, grad_scale: float = 1.0):
        lib().call("pcm_adam_step", self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                   self.exp_avg_sq.data_ptr(), self.state.data_ptr(), self.flat_param.numel(), self.lr, self.betas[0],
                   self.betas[1], self.eps, self.weight_decay, grad_scale, torch.cuda.current_stream().cuda_stream)

    def tick(self):
        """Advance the device step counter once (then `step_range` per bucket)."""
        lib().call("pcm_adam_tick", self.state.data_ptr(), self.betas[0], self.betas[1], torch.cuda.current_stream().cuda_stream)

    def step_range(self, lo: int, hi: int, grad_scale: float = 1.0):
        """Adam update of flat elements [lo, hi) with the bias corrections of the last `tick()`."""
        if hi <= lo:
            return
        o = 4 * lo
        lib().call("pcm_adam_apply", self.flat_param.data_ptr() + o, self.flat_grad.data_ptr() + o, self.exp_avg.data_ptr() + o,
                   self.exp_avg_sq.data_ptr() + o, self.state.data_ptr(), hi - lo, self.lr, self.betas[0], self.betas[1],
                   self.eps, self.weight_decay, grad_scale, torch.cuda.current_stream().cuda_stream)

    # ---- checkpointing in torch.optim.Adam's layout (Lightning stores it under `optimizer_states`) -----------------
    def state_dict(self) -> dict:
        """{"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]} with parameters numbered in
        `model.parameters()` order like torch.optim.Adam (main_final.py:737-747), sliced out of the flat buffers."""
        step = torch.tensor(float(self.state[0].item()))
        order = self._torch_order
        state = {}
        for i, p in enumerate(order):
            off, n = self._offsets[id(p)], p.numel()
            state[i] = {"step": step.clone(), "exp_avg": self.exp_avg[off:off + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[off:off + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "decoupled_weight_decay": False, "params": list(range(len(order)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        """Accepts torch.optim.Adam's state_dict (e.g. a Lightning checkpoint's `optimizer_states[0]`).  Parameters
        without an entry (never stepped: the reference's unused `post_conv`) keep zero moments."""
        order = self._torch_order
        g = sd["param_groups"][0]
        if len(g["params"]) != len(order):
            raise ValueError(f"optimizer state has {len(g['params'])} parameters, the model has {len(order)}")
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])
        self.weight_decay = float(g.get("weight_decay", 0.0))
        self.exp_avg.zero_(); self.exp_avg_sq.zero_()
        step = 0.0
        for i, pid in enumerate(g["params"]):
            st = sd["state"].get(pid)
            if st is None:
                continue
            p = order[i]
            off, n = self._offsets[id(p)], p.numel()
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, float(st["step"]))
        b1, b2 = self.betas
        self.state.copy_(torch.tensor([step, 1.0 - b1 ** step, 1.0 - b2 ** step, 0.0]))
