"""FusedAdam: the caller's `torch.optim.Adam` (AVE/traintest_adapt_ave29.py:68) as ONE kernel launch.

    optimizer = FusedAdam([{'params': adapter_params, 'lr': lr}, {'params': head_params, 'lr': lr * head_lr}],
                          weight_decay=5e-7, betas=(0.95, 0.999))

Same constructor keywords, defaults, `param_groups`, `state` keys ('step', 'exp_avg', 'exp_avg_sq') and update rule as
torch.optim.Adam (L2 weight decay added to the gradient, bias-corrected step, no amsgrad / maximize), so LR schedulers,
`GradScaler.step()` (which un-scales the .grad tensors first and skips the step on inf / nan) and `state_dict()`
checkpoints keep working.  The ~455 trainable tensors of a Swin-T adapter model are described by a device-resident
table; `step()` is a single `stgcma_adam_step` launch (stock torch: ~10 multi-tensor launches and ~1 ms of host time).
There is no fallback: parameters must be fp32 CUDA tensors on an sm_100 device.
"""
import math

import torch

from . import lib as L_


class FusedAdam(torch.optim.Optimizer):
    # torch.amp.GradScaler protocol of fused optimizers: the scaler does NOT un-scale the gradients or read found_inf on the
    # host; it hands `self.grad_scale` / `self.found_inf` (device scalars) to step(), and the kernel divides by the scale and
    # skips the update on inf / nan itself (stgcma_adam_step_amp) -- no per-step host synchronisation in the caller's loop
    # (traintest_adapt_ave29.py:165-168), which is what lets a launch-bound step run ahead / replay from CUDA graphs.
    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise L_.StgcmaError("FusedAdam: amsgrad / maximize are not used by the STG-CMA callers and are not implemented")
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0 or not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError("FusedAdam: invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        if len(self.param_groups) > L_.ADAM_MAX_GROUPS:
            raise L_.StgcmaError(f"FusedAdam: at most {L_.ADAM_MAX_GROUPS} parameter groups")
        self._table, self._key, self._n, self._numel = None, None, 0, 0
        self._skipped = None      # int32 device counter of steps skipped on inf / nan (GradScaler mode)
        self._one = self._zero = None

    def _init_state(self):
        """exp_avg / exp_avg_sq of every parameter are views into two flat buffers (allocated once)."""
        todo = [p for g in self.param_groups for p in g["params"] if p.requires_grad and "exp_avg" not in self.state[p]]
        if not todo:
            return
        dev = todo[0].device
        total = sum(p.numel() for p in todo)
        m, v = torch.zeros(total, device=dev), torch.zeros(total, device=dev)
        off = 0
        for p in todo:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise L_.StgcmaError("FusedAdam: parameters must be contiguous fp32 CUDA tensors")
            n = p.numel()
            st = self.state[p]
            st["step"] = 0
            st["exp_avg"], st["exp_avg_sq"] = m[off:off + n].view_as(p), v[off:off + n].view_as(p)
            off += n

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._init_state()
        # hyper-parameter slots: one per (param group, step count) pair -- normally one per group; a tensor that skipped
        # steps (no gradient) keeps its own bias correction, exactly like torch's per-parameter `step`
        active, key, slots = [], [], {}
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if p.grad is None:
                    continue
                gr = p.grad
                if gr.dtype != torch.float32 or not gr.is_contiguous() or gr.is_sparse:
                    raise L_.StgcmaError("FusedAdam: gradients must be dense contiguous fp32 tensors")
                st = self.state[p]
                # a torch.optim.Adam checkpoint stores `step` as a tensor: coerce, so equal counts share a slot
                st["step"] = int(st["step"]) + 1
                m, v = st["exp_avg"], st["exp_avg_sq"]
                if m.dtype != torch.float32 or v.dtype != torch.float32 or not m.is_contiguous() or not v.is_contiguous() \
                        or m.device != p.device or v.device != p.device:
                    raise L_.StgcmaError("FusedAdam: exp_avg / exp_avg_sq must be contiguous fp32 tensors on the parameter's device")
                slot = slots.setdefault((gi, st["step"]), len(slots))
                active.append((slot, p, gr))
                # the moment buffers are part of the key: load_state_dict() replaces them
                key.append((p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), slot))
        if not active:
            return loss
        if len(slots) > L_.ADAM_MAX_GROUPS:
            raise L_.StgcmaError(f"FusedAdam: more than {L_.ADAM_MAX_GROUPS} distinct (group, step) combinations")
        key = tuple(key)
        if key != self._key:          # pointers changed (first step, new gradient bucket): rebuild the device table
            raw = b"".join(L_.adam_entry(p, gr, self.state[p]["exp_avg"], self.state[p]["exp_avg_sq"], slot)
                           for slot, p, gr in active)
            self._table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(active[0][1].device)
            self._key, self._n = key, len(active)
            self._numel = sum(p.numel() for _, p, _ in active)
        groups, steps = [None] * len(slots), [1] * len(slots)
        for (gi, t), slot in slots.items():
            g = self.param_groups[gi]
            b1, b2 = g["betas"]
            groups[slot] = dict(lr=g["lr"], weight_decay=g["weight_decay"], beta1=b1, beta2=b2, eps=g["eps"],
                                bias_corr1=1.0 - b1 ** t, bias_corr2_sqrt=math.sqrt(1.0 - b2 ** t))
            steps[slot] = t
        amp_scale, found_inf = getattr(self, "grad_scale", None), getattr(self, "found_inf", None)
        if isinstance(amp_scale, torch.Tensor) or isinstance(found_inf, torch.Tensor) or self._skipped is not None:
            dev = active[0][1].device
            if self._skipped is None:
                self._skipped = torch.zeros(1, dtype=torch.int32, device=dev)
                self._one, self._zero = torch.ones(1, device=dev), torch.zeros(1, device=dev)
            sc = amp_scale if isinstance(amp_scale, torch.Tensor) else self._one
            fi = found_inf if isinstance(found_inf, torch.Tensor) else self._zero
            sc = sc.to(device=dev, dtype=torch.float32).reshape(1)
            fi = fi.to(device=dev, dtype=torch.float32).reshape(1)
            if grad_scale != 1.0:
                sc = sc / grad_scale
            L_.adam_step_amp(self._table, self._n, groups, steps, sc, fi, self._skipped, numel=self._numel)
        else:
            L_.adam_step(self._table, self._n, groups, grad_scale, numel=self._numel)
        return loss

    def skipped_steps(self):
        """Steps the GradScaler path skipped on inf / nan gradients (one host read)."""
        return 0 if self._skipped is None else int(self._skipped.item())

    def state_dict(self):
        """`step` counts APPLIED updates, as torch.optim.Adam's does (the host counter also counts the skipped ones)."""
        sd = super().state_dict()
        k = self.skipped_steps()
        if k:
            for st in sd["state"].values():
                if "step" in st:
                    st["step"] = int(st["step"]) - k
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        if self._skipped is not None:
            self._skipped.zero_()
