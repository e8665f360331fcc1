"""Fused optimiser step: the reference's update sequence

    scaler.unscale_(optimizer)                                                   (--fp16 only)
    grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=5.0)
    optimizer.step()  /  scaler.step(optimizer); scaler.update()                 (skipped when the norm is not finite)

with `optimizer = optim.AdamW(params, lr, weight_decay, amsgrad=True)` (train_ContSep.py:233,402-419;
train_ContExt.py:372-389) as ONE C-ABI call (`cse_optim_step`: three launches, no host synchronisation).

`AdamW` keeps torch.optim.AdamW's constructor, `param_groups` (LR schedulers keep working) and `state_dict()` layout
(`step`, `exp_avg`, `exp_avg_sq`, `max_exp_avg_sq` per parameter), so optimiser checkpoints are interchangeable with
the reference's.  No CPU fallback: parameters and gradients must be CUDA float32 tensors.
"""
import ctypes as C

import torch

from . import _lib
from .runtime import current_stream

_STATE_FLOATS = 16


class AdamW(torch.optim.Optimizer):
    """torch.optim.AdamW(params, lr, betas, eps, weight_decay, amsgrad) with the gradient clipping and the
    GradScaler bookkeeping of the reference's training loop folded into `step`.

        opt = AdamW(model.parameters(), lr=1e-4, weight_decay=1e-6, amsgrad=True)
        loss.backward();                         grad_norm = opt.step(max_norm=5.0)      # --bf16 / fp32
        opt.scale(loss).backward();              grad_norm = opt.step(max_norm=5.0)      # --fp16 (init_scale given)

    `step` returns the total gradient norm as a 0-dim CUDA tensor (what clip_grad_norm_ returns) without
    synchronising; `opt.found_inf` / `opt.get_scale()` read the device state (they synchronise)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, *,
                 init_scale=None, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self._use_scaler = init_scale is not None
        self._scaler = (float(init_scale or 1.0), float(growth_factor), float(backoff_factor), int(growth_interval))
        self._dev = {}      # group index -> dict(state, table, key, n_chunks, partial, keep)

    # ---- device state ------------------------------------------------------------------------------------
    def _group_dev(self, gi, device):
        d = self._dev.get(gi)
        if d is None:
            st = torch.zeros(_STATE_FLOATS, dtype=torch.float32, device=device)
            st[2] = self._scaler[0]
            d = self._dev[gi] = dict(state=st, key=None, gptrs=None, table=None, n_chunks=0, partial=None)
        return d

    def _moments(self, p, amsgrad):
        s = self.state[p]
        if "exp_avg" not in s:
            s["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            s["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            if amsgrad:
                s["max_exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return s

    @staticmethod
    def _check(t, what):
        if not t.is_cuda:
            raise _lib.CseError(f"{what} is on {t.device}: the CUDA path has no CPU fallback")
        if t.dtype != torch.float32 or not t.is_contiguous() or t.is_sparse:
            raise _lib.CseError(f"{what} must be a dense contiguous float32 tensor (got {t.dtype})")

    def _table(self, d, ps, amsgrad):
        """Chunk table on the device.  Parameters and moments keep their storage from step to step; autograd hands
        out new `.grad` storages (zero_grad(set_to_none=True), DDP bucket copies), so usually only the gradient
        pointers change: they are rewritten in a rotating pinned staging copy and re-uploaded (one small H2D)."""
        static_key = tuple(p.data_ptr() for p in ps) + (amsgrad,)
        gptrs = [p.grad.data_ptr() for p in ps]
        n = len(ps)
        if d["key"] != static_key:
            for p in ps:
                self._check(p, "parameter")
            numel = (C.c_longlong * n)(*[p.numel() for p in ps])
            arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])  # noqa: E731
            sts = [self._moments(p, amsgrad) for p in ps]
            n_chunks = _lib.load().cse_optim_chunk_count(n, numel)
            # four rotating staging copies for eager steps + one owned by a captured step (CUDA graph)
            hosts = [torch.empty(n_chunks * 48, dtype=torch.uint8).pin_memory() for _ in range(5)]
            garr = (C.c_void_p * n)(*gptrs)
            for h in hosts:
                _lib.call("cse_optim_table_fill", n, numel, arr(ps), garr, arr([s["exp_avg"] for s in sts]),
                          arr([s["exp_avg_sq"] for s in sts]),
                          arr([s["max_exp_avg_sq"] for s in sts]) if amsgrad else None,
                          C.c_void_p(h.data_ptr()), h.numel())
            dev = ps[0].device
            for t in (ps[0].grad, ps[-1].grad):
                self._check(t, "gradient")
            d.update(key=static_key, gptrs=None, numel=numel, n_chunks=n_chunks, hosts=hosts, turn=0,
                     events=[None] * 4, graph_gptrs=None,
                     table=torch.empty(n_chunks * 48, dtype=torch.uint8, device=dev),
                     graph_table=torch.empty(n_chunks * 48, dtype=torch.uint8, device=dev),
                     partial=torch.empty(n_chunks, dtype=torch.float32, device=dev))
        if torch.cuda.is_current_stream_capturing():
            # A captured step (runtime.GraphedStep): its gradients live at fixed addresses in the graph's pool.  The
            # upload becomes a node of the graph and re-reads ITS OWN pinned copy and device table on every replay, so
            # eager steps of the same optimiser (which rotate the other four copies) cannot disturb it.
            if d["graph_gptrs"] is not None and d["graph_gptrs"] != gptrs:
                raise _lib.CseError("AdamW: one captured step per optimiser (gradient addresses differ from the "
                                    "captured ones)")
            h = d["hosts"][4]
            _lib.call("cse_optim_table_set_grads", n, d["numel"], (C.c_void_p * n)(*gptrs), C.c_void_p(h.data_ptr()),
                      h.numel())
            d["graph_table"].copy_(h, non_blocking=True)
            d["graph_gptrs"] = gptrs
            return d["graph_table"]
        if d["gptrs"] != gptrs:
            slot = d["turn"] & 3
            d["turn"] += 1
            h = d["hosts"][slot]
            if d["events"][slot] is not None:      # its previous upload (four refreshes ago) has been consumed
                d["events"][slot].synchronize()
            _lib.call("cse_optim_table_set_grads", n, d["numel"], (C.c_void_p * n)(*gptrs), C.c_void_p(h.data_ptr()),
                      h.numel())
            d["table"].copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            d["events"][slot] = ev
            d["gptrs"] = gptrs
        return d["table"]

    # ---- the update --------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None, max_norm=None, write_back_grads=False):
        """One update.  max_norm: clip_grad_norm_'s max_norm (None: no clipping).  Returns the gradient norm."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        norm = None
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            if not ps[0].is_cuda:
                raise _lib.CseError(f"parameter is on {ps[0].device}: the CUDA path has no CPU fallback")
            d = self._group_dev(gi, ps[0].device)
            table = self._table(d, ps, bool(group["amsgrad"]))
            _, gf, bf, gint = self._scaler
            b1, b2 = group["betas"]
            _lib.call("cse_optim_step", _lib.ptr(table), d["n_chunks"], float(group["lr"]), float(b1), float(b2),
                      float(group["eps"]), float(group["weight_decay"]), int(bool(group["amsgrad"])),
                      float(max_norm) if max_norm is not None else 0.0, int(self._use_scaler), gf, bf, gint,
                      int(bool(write_back_grads)), _lib.ptr(d["state"]), _lib.ptr(d["partial"]),
                      C.c_void_p(current_stream(ps[0].device)))
            norm = d["state"][4]
        return norm if closure is None else loss

    # ---- GradScaler surface (train_ContSep.py:175,397,405-410,446) ------------------------------------------
    def scale(self, loss):
        """scaler.scale(loss): multiplies by the device-resident scale (no synchronisation)."""
        if not self._use_scaler:
            return loss
        d = self._group_dev(0, loss.device)
        return loss * d["state"][2]

    def get_scale(self):
        d = self._dev.get(0)
        return float(d["state"][2].item()) if d is not None else self._scaler[0]

    @property
    def found_inf(self):
        """True when the last `step` met non-finite gradients and skipped the update."""
        return any(bool(d["state"].view(torch.int32)[6].item()) for d in self._dev.values())

    def steps_applied(self, group=0):
        d = self._dev.get(group)
        return int(d["state"].view(torch.float64)[0].item()) if d is not None else 0

    # ---- checkpoint layout of torch.optim.AdamW -------------------------------------------------------------
    def state_dict(self):
        for gi, group in enumerate(self.param_groups):
            n = self.steps_applied(gi)
            for p in group["params"]:
                if p in self.state and "exp_avg" in self.state[p]:
                    self.state[p]["step"] = torch.tensor(float(n))
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._dev.clear()
        for gi, group in enumerate(self.param_groups):
            steps = [float(self.state[p]["step"]) for p in group["params"] if p in self.state and "step" in self.state[p]]
            dev = next((p.device for p in group["params"] if p.is_cuda), None)
            if steps and dev is not None:
                self._group_dev(gi, dev)["state"].view(torch.float64)[0] = max(steps)
