"""Flat-buffer RMSprop with fused global-norm clipping (two kernel launches per step).

Drop-in for `torch.optim.RMSprop(model.parameters(), lr, momentum, eps, alpha)` as the
reference builds it (/root/reference/torchbeast/monobeast.py:388-394,
polybeast_learner.py:472-478) for a FlatParamModule: same constructor arguments,
`param_groups[0]["lr"]` is honoured so `torch.optim.lr_scheduler.LambdaLR` works unchanged,
`state_dict()` carries square_avg (and momentum_buffer) per parameter like torch's.
"""
import torch

from torchbeast_b200 import _lib


class RMSprop(torch.optim.Optimizer):
    def __init__(self, model, lr=0.01, alpha=0.99, eps=1e-8, weight_decay=0, momentum=0, centered=False):
        if weight_decay != 0 or centered:
            raise _lib.TorchBeastB200Error("torchbeast_b200.optim.RMSprop: weight_decay/centered are not supported")
        if not hasattr(model, "flat_params"):
            raise _lib.TorchBeastB200Error("torchbeast_b200.optim.RMSprop needs a FlatParamModule (e.g. AtariNet)")
        self.model = model
        defaults = dict(lr=lr, alpha=alpha, eps=eps, momentum=momentum, weight_decay=0, centered=False)
        super().__init__(list(model.parameters()), defaults)
        flat = model.flat_params
        self.square_avg = torch.zeros_like(flat)
        self.momentum_buffer = torch.zeros_like(flat) if momentum != 0 else None
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=flat.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=flat.device)
        self._steps = 0
        # learning rate as a device scalar: lets a captured CUDA graph see scheduler updates
        self._lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=flat.device)
        self.lr_from_device = False
        self._point_state_at_flat()

    def _point_state_at_flat(self, step=None):
        """Per-parameter views of the flat buffers so state_dict() looks like torch.optim.RMSprop's."""
        for p, off, n, shape in self.model._views:
            st = self.state[p]
            if step is not None or "step" not in st:
                st["step"] = torch.tensor(float(step or 0.0))
            st["square_avg"] = self.square_avg[off:off + n].view(shape)
            if self.momentum_buffer is not None:
                st["momentum_buffer"] = self.momentum_buffer[off:off + n].view(shape)

    def load_state_dict(self, state_dict):
        """Resume like the reference does (`optimizer.load_state_dict(checkpoint["optimizer_state_dict"])`,
        polybeast_learner.py:535-548): torch replaces the per-parameter state tensors with copies, so the loaded
        values are copied INTO the flat buffers the fused kernel reads and the per-parameter entries are re-pointed
        at them.  Accepts state_dicts of torch.optim.RMSprop (same per-parameter keys) and of this class."""
        self._loading = True  # torch calls __setstate__ with the loaded copies: keep them until they are in the flat buffers
        try:
            super().load_state_dict(state_dict)
        finally:
            self._loading = False
        steps = 0
        with torch.no_grad():
            for p, off, n, shape in self.model._views:
                st = self.state.get(p, {})
                if "square_avg" in st:
                    self.square_avg[off:off + n].copy_(st["square_avg"].reshape(-1))
                if self.momentum_buffer is not None and st.get("momentum_buffer") is not None:
                    self.momentum_buffer[off:off + n].copy_(st["momentum_buffer"].reshape(-1))
                if "step" in st:
                    steps = max(steps, int(float(st["step"])))
        self._steps = steps
        self._point_state_at_flat(step=steps)
        self._lr_dev.fill_(float(self.param_groups[0]["lr"]))

    def __setstate__(self, state):
        super().__setstate__(state)
        if hasattr(self, "model") and not getattr(self, "_loading", False):
            self._point_state_at_flat()

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None):
        """clip_grad_norm_(max_grad_norm) (skipped when None) + RMSprop on the flat buffers."""
        if closure is not None:
            raise _lib.TorchBeastB200Error("closure is not supported")
        model = self.model
        flat, grad = model.flat_params, model.attach_grads()
        group = self.param_groups[0]
        lib, p = _lib.lib(), _lib.ptr
        st = _lib.stream_ptr()
        _lib.check(lib.tb_grad_sumsq_f32(p(grad), grad.numel(), p(self._sumsq), p(_lib.workspace()), st),
                   "tb_grad_sumsq_f32")
        _lib.check(
            lib.tb_clip_rmsprop_step_f32(
                p(flat), p(grad), p(self.square_avg), p(self.momentum_buffer), flat.numel(), p(self._sumsq),
                -1.0 if max_grad_norm is None else float(max_grad_norm),
                p(self._lr_dev) if self.lr_from_device else None, float(group["lr"]),
                float(group["alpha"]), float(group["eps"]), float(group["momentum"]), p(self.grad_norm), st),
            "tb_clip_rmsprop_step_f32")
        self._steps += 1
        return None

    def state_dict(self):
        for st in self.state.values():  # torch.optim.RMSprop keeps a per-parameter step count in its state_dict
            st["step"] = torch.tensor(float(self._steps))
        return super().state_dict()

    def sync_lr_to_device(self):
        """Publish param_groups[0]['lr'] to the device scalar a captured graph reads (async fill)."""
        self._lr_dev.fill_(float(self.param_groups[0]["lr"]))
