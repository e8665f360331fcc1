"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's optimiser step.

The reference calls stock PyTorch directly (train_ContSep.py:233,402-419):
    optimizer = optim.AdamW(params, lr, weight_decay, amsgrad=True)
    scaler.unscale_(optimizer); grad_norm = clip_grad_norm_(model.parameters(), max_norm=5.0)
    scaler.step(optimizer); scaler.update()        /  optimizer.step() unless the norm is not finite
so the pin is torch itself: tests/test_optim.py checks this restatement against torch.optim.AdamW +
torch.nn.utils.clip_grad_norm_ (+ torch.amp.GradScaler's documented update rule) on the CPU, and the CUDA kernels
against this restatement.  Formulas: torch/optim/adam.py:_single_tensor_adam (decoupled weight decay, amsgrad),
torch/nn/utils/clip_grad.py, torch/amp/grad_scaler.py:_amp_update_scale_.  Never imported by the product path."""
import math

import numpy as np


class AdamWOracle:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False,
                 init_scale=None, growth_factor=2.0, backoff_factor=0.5, growth_interval=2000):
        self.p = [np.array(p, dtype=np.float32) for p in params]
        self.m = [np.zeros_like(p) for p in self.p]
        self.v = [np.zeros_like(p) for p in self.p]
        self.vmax = [np.zeros_like(p) for p in self.p]
        self.lr, self.betas, self.eps, self.wd, self.amsgrad = lr, betas, eps, weight_decay, amsgrad
        self.step_count = 0
        self.use_scaler = init_scale is not None
        self.scale = float(init_scale or 1.0)
        self.growth = (growth_factor, backoff_factor, growth_interval)
        self.tracker = 0
        self.found_inf = False

    def step(self, grads, max_norm=None):
        """grads: gradients of the SCALED loss.  Returns the norm of the unscaled gradients."""
        f32 = np.float32
        inv = 1.0 / self.scale
        sq = sum(float(np.sum(np.square(g.astype(np.float64)))) for g in grads)
        total = math.sqrt(sq) * inv if math.isfinite(sq) else float("nan")
        self.found_inf = not math.isfinite(total)
        if not self.found_inf:
            coef = min(1.0, max_norm / (total + 1e-6)) if max_norm else 1.0
            c = f32(coef * inv)
            self.step_count += 1
            b1, b2 = self.betas
            step_size = f32(self.lr / (1.0 - b1 ** self.step_count))
            bc2_sqrt = f32(math.sqrt(1.0 - b2 ** self.step_count))
            for i, g in enumerate(grads):
                g = g.astype(f32) * c
                self.p[i] = self.p[i] * f32(1.0 - self.lr * self.wd)
                self.m[i] = self.m[i] + f32(1.0 - b1) * (g - self.m[i])
                self.v[i] = self.v[i] * f32(b2) + f32(1.0 - b2) * g * g
                d = self.v[i]
                if self.amsgrad:
                    self.vmax[i] = np.maximum(self.vmax[i], self.v[i])
                    d = self.vmax[i]
                denom = np.sqrt(d) / bc2_sqrt + f32(self.eps)
                self.p[i] = self.p[i] - step_size * (self.m[i] / denom)
        if self.use_scaler:
            gf, bf, gi = self.growth
            if self.found_inf:
                self.scale *= bf
                self.tracker = 0
            else:
                self.tracker += 1
                if self.tracker == gi:
                    self.scale *= gf
                    self.tracker = 0
        return total
