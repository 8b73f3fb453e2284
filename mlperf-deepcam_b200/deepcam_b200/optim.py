"""Fused multi-tensor Adam / AdamW / LAMB for the deepcam_b200 modules.

Drop-in for `torch.optim.Adam(params, lr, betas, eps, weight_decay)` / `torch.optim.AdamW` as the reference constructs
them (TR:213-220): same constructor arguments, same `param_groups`, same per-parameter state (`step`, `exp_avg`,
`exp_avg_sq`), so `optimizer.state_dict()` checkpoints (TR:519) are interchangeable with the stock optimizers.  The
update of every parameter of a group runs in ONE sm_100a kernel launch (csrc/optim.cu) instead of ~10 foreach kernels.
`torch.optim.Adam` itself keeps working with the modules unchanged; this class is an optional accelerator.
"""
import ctypes
import math

import torch

from . import _lib, ops


def _mark_updated(plist):
    """The kernels write the parameters through raw pointers, which autograd's version counters do not see: bump them so
    every consumer keyed on `param._version` (the packed bf16 weight copies of deepcam_b200.backend, torch's own
    saved-tensor checks) notices the update exactly as it would after `torch.optim.Adam.step()`."""
    torch._C._increment_version(list(plist))


class _FusedAdamBase(torch.optim.Optimizer):
    _adamw = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("deepcam_b200 fused Adam: amsgrad is not supported")
        if lr < 0.0 or eps < 0.0 or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False))
        self._tables = {}

    def _table(self, gi, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in plist)
        ent = self._tables.get(gi)
        if ent is not None and ent[0] == key:
            return ent[1]
        arr = (_lib.dc_adam_job * len(plist))()
        start = 0
        for i, p in enumerate(plist):
            st = self.state[p]
            n = p.numel()
            nb = max(1, min(256, (n + 4095) // 4096))
            j = arr[i]
            j.p, j.g, j.m, j.v = p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            j.numel, j.block_start, j.n_blocks = n, start, nb
            start += nb
        dev = plist[0].device
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
        self._tables[gi] = (key, (table, len(plist), start))
        return self._tables[gi][1]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("deepcam_b200 fused Adam needs fp32 CUDA parameters and gradients (no CPU fallback)")
                if p.grad.is_sparse or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("deepcam_b200 fused Adam needs dense contiguous parameters and gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = {float(self.state[p]["step"]) for p in plist}
            if len(steps) != 1:
                raise RuntimeError("deepcam_b200 fused Adam: parameters of one group must share the step count")
            t = steps.pop() + 1.0
            for p in plist:
                self.state[p]["step"] += 1
            b1, b2 = group["betas"]
            table, njobs, blocks = self._table(gi, plist)
            _lib.check(_lib.load().dc_adam_step_multi(
                ctypes.c_void_p(table.data_ptr()), njobs, blocks, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                float(group["weight_decay"]), 1.0 - math.pow(b1, t), 1.0 - math.pow(b2, t), int(self._adamw),
                ops._stream()), "dc_adam_step_multi")
            _mark_updated(plist)
        return loss


class FusedAdam(_FusedAdamBase):
    """torch.optim.Adam semantics (weight decay is L2: added to the gradient)."""
    _adamw = False


class FusedAdamW(_FusedAdamBase):
    """torch.optim.AdamW semantics (decoupled weight decay)."""
    _adamw = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False):
        super().__init__(params, lr, betas, eps, weight_decay, amsgrad)


class FusedLAMB(torch.optim.Optimizer):
    """Drop-in for `apex.optimizers.FusedLAMB(params, lr, eps, weight_decay)` as the reference constructs it for
    `--optimizer LAMB` (TR:217-218; apex is absent from this image, so the reference cannot take that branch here).
    Same constructor arguments and defaults as apex (bias_correction, betas, eps=1e-6, weight_decay=0.01, adam_w_mode,
    grad_averaging, max_grad_norm=1.0, use_nvlamb), same state layout (`exp_avg`, `exp_avg_sq` per parameter, `step` per
    group).  One step = three sm_100a launches for all parameters (csrc/optim.cu: gradient norm, moments + update + per-tensor
    norms, trust-ratio update).  Like apex, the step overwrites `.grad` with the update."""

    def __init__(self, params, lr=1e-3, bias_correction=True, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.01, amsgrad=False,
                 adam_w_mode=True, grad_averaging=True, set_grad_none=True, max_grad_norm=1.0, use_nvlamb=False):
        if amsgrad:
            raise RuntimeError("FusedLAMB does not support the AMSGrad variant.")
        super().__init__(params, dict(lr=lr, bias_correction=bias_correction, betas=betas, eps=eps, weight_decay=weight_decay,
                                      grad_averaging=grad_averaging, max_grad_norm=max_grad_norm))
        self.adam_w_mode = 1 if adam_w_mode else 0
        self.set_grad_none = set_grad_none
        self.use_nvlamb = use_nvlamb
        self._tables = {}

    def zero_grad(self, set_to_none=None):
        super().zero_grad(set_to_none=self.set_grad_none if set_to_none is None else set_to_none)

    _table = _FusedAdamBase._table

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        # apex clips by the norm over ALL groups; with one group (the reference's case) that is the group's own norm
        if len([g for g in self.param_groups if any(p.grad is not None for p in g["params"])]) > 1:
            raise NotImplementedError("deepcam_b200 FusedLAMB: a single parameter group is supported (as in TR:217-218)")
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("deepcam_b200 FusedLAMB needs fp32 CUDA parameters and gradients (no CPU fallback)")
                if p.grad.is_sparse or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("deepcam_b200 FusedLAMB needs dense contiguous parameters and gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
            group["step"] = group.get("step", 0) + 1
            t = group["step"]
            b1, b2 = group["betas"]
            bc1 = 1.0 - math.pow(b1, t) if group["bias_correction"] else 1.0
            bc2 = 1.0 - math.pow(b2, t) if group["bias_correction"] else 1.0
            table, njobs, blocks = self._table(gi, plist)
            norms = torch.empty(1 + 2 * njobs, dtype=torch.float64, device=plist[0].device)
            _lib.check(_lib.load().dc_lamb_step_multi(
                ctypes.c_void_p(table.data_ptr()), njobs, blocks, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                float(group["weight_decay"]), bc1, bc2, self.adam_w_mode, int(bool(group["grad_averaging"])),
                float(group["max_grad_norm"] or 0.0), int(bool(self.use_nvlamb)), ctypes.c_void_p(norms.data_ptr()),
                ops._stream()), "dc_lamb_step_multi")
            _mark_updated(plist)
            _mark_updated([p.grad for p in plist])      # like apex, the step leaves the update in .grad
        return loss


class FusedLARS(torch.optim.Optimizer):
    """LARS (You et al. 2017) for the local-batch sweep of BASELINE.json configs[4].  The reference ships no LARS (its
    layer-wise optimizer is apex FusedLAMB, see FusedLAMB above), so the update rule is the commonly used one and its parity
    is pinned only against a plain-torch restatement ("parity unpinned"):
        local_lr = trust_coefficient * ||w|| / (||g|| + weight_decay * ||w|| + eps)     per tensor (1 if a norm is 0)
        buf = momentum * buf + lr * local_lr * (g + weight_decay * w);   w -= buf
    State: `momentum_buffer` per parameter.  One step = two launches for all parameters (csrc/optim.cu)."""

    def __init__(self, params, lr=1e-3, momentum=0.9, weight_decay=0.0, trust_coefficient=0.001, eps=1e-8):
        if lr < 0.0 or momentum < 0.0 or weight_decay < 0.0 or trust_coefficient <= 0.0:
            raise ValueError("invalid LARS hyper-parameters")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, trust_coefficient=trust_coefficient, eps=eps))
        self._tables = {}

    def _table(self, gi, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["momentum_buffer"].data_ptr()) for p in plist)
        ent = self._tables.get(gi)
        if ent is not None and ent[0] == key:
            return ent[1]
        arr = (_lib.dc_adam_job * len(plist))()
        start = 0
        for i, p in enumerate(plist):
            n = p.numel()
            nb = max(1, min(256, (n + 4095) // 4096))
            j = arr[i]
            j.p, j.g, j.m, j.v = p.data_ptr(), p.grad.data_ptr(), self.state[p]["momentum_buffer"].data_ptr(), None
            j.numel, j.block_start, j.n_blocks = n, start, nb
            start += nb
        table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(plist[0].device)
        self._tables[gi] = (key, (table, len(plist), start))
        return self._tables[gi][1]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("deepcam_b200 FusedLARS needs fp32 CUDA parameters and gradients (no CPU fallback)")
                if p.grad.is_sparse or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("deepcam_b200 FusedLARS needs dense contiguous parameters and gradients")
                if "momentum_buffer" not in self.state[p]:
                    self.state[p]["momentum_buffer"] = torch.zeros_like(p)
            table, njobs, blocks = self._table(gi, plist)
            norms = torch.empty(2 * njobs, dtype=torch.float64, device=plist[0].device)
            _lib.check(_lib.load().dc_lars_step_multi(
                ctypes.c_void_p(table.data_ptr()), njobs, blocks, float(group["lr"]), float(group["momentum"]),
                float(group["weight_decay"]), float(group["trust_coefficient"]), float(group["eps"]),
                ctypes.c_void_p(norms.data_ptr()), ops._stream()), "dc_lars_step_multi")
            _mark_updated(plist)
        return loss
