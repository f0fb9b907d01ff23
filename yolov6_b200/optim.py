"""Fused SGD-nesterov + weight decay + EMA over the flat parameter / gradient buffers.

Replaces, for a model whose state lives in a `FlatState` (flat.py), what the reference does with
`torch.optim.SGD(momentum, nesterov=True)` over the three parameter groups of `build_optimizer`
(yolov6/solver/build.py:10-33: BatchNorm weights / conv weights with weight decay / biases) followed by
`ModelEMA.update` (yolov6/utils/ema.py:28-37: decay = d * (1 - exp(-updates / 2000)), every floating-point entry of
the state_dict incl. BatchNorm running statistics) -- ~600 small kernels and a Python walk over the state_dict per
step -- with ONE kernel launch (`yv6_sgd_ema_step`).  Learning rates / momentum / decay are read from device memory,
so a CUDA graph that contains the launch follows the schedule (`set_lr`, warm-up of core/engine.py:360-376).
"""
import math

import numpy as np
import torch

from . import _lib


class FusedSGDEMA:
    def __init__(self, model, lr=0.01, momentum=0.937, weight_decay=5e-4, ema_decay=0.9999, ema=True, ema_updates=0):
        self.model = model
        self.engine = model.train_engine()
        self.flat = fl = self.engine.flat
        self.dev = fl.device
        self.lib, self.h = _lib.lib(), _lib.handle(self.dev.index or 0)
        self.lrs = [float(lr)] * 3                   # groups: BN weights, conv weights, biases (build.py:12-19)
        self.momentum, self.weight_decay = float(momentum), float(weight_decay)
        self.ema_decay_base, self.updates = float(ema_decay), int(ema_updates)
        self.mom = torch.zeros(fl.total, dtype=torch.float32, device=self.dev)
        self.ema = fl.pflat.clone() if ema else None          # ModelEMA starts as a copy of the model (ema.py:22)
        self.hyper = torch.zeros(8, dtype=torch.float32, device=self.dev)
        self.steps = 0
        self.grad_scale = 1.0

    def set_lr(self, lrs):
        self.lrs = [float(v) for v in (lrs if isinstance(lrs, (list, tuple)) else [lrs] * 3)]

    def warmup(self, curr_step, warmup_steps, lr_target, warmup_bias_lr=0.1, warmup_momentum=0.8, momentum=0.937):
        """The per-step warm-up of Trainer.update_optimizer (core/engine.py:360-372)."""
        if curr_step <= warmup_steps:
            self.lrs = [float(np.interp(curr_step, [0, warmup_steps], [warmup_bias_lr if k == 2 else 0.0, lr_target])) for k in range(3)]
            self.momentum = float(np.interp(curr_step, [0, warmup_steps], [warmup_momentum, momentum]))

    def upload_hyper(self):
        """Host -> device copy of this step's hyper-parameters (outside any captured graph)."""
        self.updates += 1
        d = self.ema_decay_base * (1 - math.exp(-self.updates / 2000))       # ema.py:23
        vals = [*self.lrs, self.momentum, self.weight_decay, d, 1.0 if self.steps == 0 else 0.0, self.grad_scale]
        self.hyper.copy_(torch.tensor(vals, dtype=torch.float32))
        self.steps += 1

    def launch(self, stream=None):
        """The kernel launch alone (capturable); `upload_hyper` must have run for this step."""
        fl = self.flat
        _lib.check(self.lib.yv6_sgd_ema_step(self.h, fl.pflat.data_ptr(), fl.gflat.data_ptr(), self.mom.data_ptr(),
                                             self.ema.data_ptr() if self.ema is not None else 0, fl.group.data_ptr(), fl.total,
                                             self.hyper.data_ptr(), _lib.stream_ptr(stream)))

    def step(self):
        self.upload_hyper()
        self.launch()
        self.model.mark_weights_changed()

    def ema_state_dict(self):
        """state_dict of the averaged model (what ModelEMA.ema.state_dict() holds): float entries from the EMA buffer,
        integer counters from the live model."""
        fl = self.flat
        src = self.ema if self.ema is not None else fl.pflat
        out = {}
        for k, v in self.model.state_dict().items():
            if k in fl.slots:
                o, n, shape = fl.slots[k]
                out[k] = src[o:o + n].view(shape).clone()
            else:
                out[k] = v.clone()
        return out
