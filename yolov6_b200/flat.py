"""Flat fp32 storage of a Model's state: one buffer for every parameter and float buffer, one for the gradients.

The reference keeps ~600 separate tensors and walks them in Python every step: autocast casts each weight,
`loss.backward()` accumulates each `.grad`, DDP copies gradients into buckets (core/engine.py:456-468),
torch.optim.SGD and ModelEMA.update loop over the state (solver/build.py:10-33, utils/ema.py:28-37).  Here all
of them are *views* of two flat buffers, laid out in the order in which the backward pass finishes them, so that

  * the gradient all-reduce is one NCCL call per contiguous bucket of the flat gradient buffer (dist.py),
  * SGD + weight decay + EMA are one kernel over the flat buffers (optim.py / yv6_sgd_ema_step),
  * the per-step weight repack (fp32 -> bf16 KRSC, dgrad layouts) and the gradient unpack are one table-driven
    launch each (train.py / yv6_xform).

`nn.Parameter.data` / BatchNorm buffers are re-pointed at the views; `state_dict()`, `load_state_dict()`,
torch optimizers, `ModelEMA` (deepcopy) and DDP keep working on them unchanged.
"""
import torch
import torch.nn as nn

GROUP_BNW, GROUP_W, GROUP_B, GROUP_EMA_ONLY, GROUP_PAD = 0, 1, 2, 3, 255


def _align4(n):
    return (n + 3) // 4 * 4


class FlatState:
    def __init__(self, model, order=()):
        """order: names of trainable parameters in backward-completion order (they are laid out first, in that
        order); every other parameter / float buffer follows."""
        named_p = dict(model.named_parameters())
        dev = next(iter(named_p.values())).device
        self.device = dev
        # parameter groups of build_optimizer (solver/build.py:12-19): bias of any module -> g_b, BatchNorm weight ->
        # g_bnw, any other `weight` -> g_w; everything else (BottleRep.alpha, detect.proj) is not optimised
        group_of = {}
        for mname, mod in model.named_modules():
            pre = mname + "." if mname else ""
            if isinstance(getattr(mod, "bias", None), nn.Parameter):
                group_of[pre + "bias"] = GROUP_B
            if isinstance(mod, nn.BatchNorm2d):
                group_of[pre + "weight"] = GROUP_BNW
            elif isinstance(getattr(mod, "weight", None), nn.Parameter):
                group_of[pre + "weight"] = GROUP_W
        names = [n for n in order if n in named_p and named_p[n].requires_grad]
        seen = set(names)
        names += [n for n, p in named_p.items() if p.requires_grad and n not in seen]
        self.n_train_names = len(names)
        rest = [n for n, p in named_p.items() if not p.requires_grad]
        fbuf = [(n, b) for n, b in model.named_buffers() if b.dtype.is_floating_point]
        ibuf = [(n, b) for n, b in model.named_buffers() if not b.dtype.is_floating_point]
        self.slots = {}          # name -> (offset, numel, shape)
        off = 0
        for n in names:
            p = named_p[n]
            self.slots[n] = (off, p.numel(), tuple(p.shape))
            off += _align4(p.numel())
        self.n_train = off       # gradients exist for [0, n_train)
        for n in rest:
            p = named_p[n]
            self.slots[n] = (off, p.numel(), tuple(p.shape))
            off += _align4(p.numel())
        for n, b in fbuf:
            self.slots[n] = (off, b.numel(), tuple(b.shape))
            off += _align4(b.numel())
        self.total = off
        self.pflat = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.gflat = torch.zeros(self.n_train, dtype=torch.float32, device=dev)
        group = torch.full((self.total // 4,), GROUP_PAD, dtype=torch.uint8)
        with torch.no_grad():
            for n in names + rest:
                p = named_p[n]
                o, k, shape = self.slots[n]
                view = self.pflat[o:o + k].view(shape)
                view.copy_(p.data)
                p.data = view
                g = group_of.get(n, GROUP_EMA_ONLY) if p.requires_grad else GROUP_EMA_ONLY
                group[o // 4:(o + _align4(k)) // 4] = g
            mods = dict(model.named_modules())
            for n, b in fbuf:
                o, k, shape = self.slots[n]
                view = self.pflat[o:o + k].view(shape)
                view.copy_(b)
                mname, _, leaf = n.rpartition(".")
                mods[mname]._buffers[leaf] = view
                group[o // 4:(o + _align4(k)) // 4] = GROUP_EMA_ONLY
            # integer buffers (BatchNorm.num_batches_tracked): one flat int64 tensor, bumped by one add per step
            self.iflat = torch.zeros(max(len(ibuf), 1), dtype=torch.int64, device=dev)
            for i, (n, b) in enumerate(ibuf):
                self.iflat[i] = b.to(dev)
                mname, _, leaf = n.rpartition(".")
                mods[mname]._buffers[leaf] = self.iflat[i]
        self.group = group.to(dev)
        self.names = names
        self._probe = named_p[names[0]] if names else None
        self._probe_ptr = self._probe.data_ptr() if names else 0

    # ------------------------------------------------------------------ views
    def param_view(self, name):
        o, k, shape = self.slots[name]
        return self.pflat[o:o + k].view(shape)

    def grad_view(self, name):
        o, k, shape = self.slots[name]
        return self.gflat[o:o + k].view(shape)

    def ptr(self, name):
        """Device address of a parameter / buffer inside the flat buffer."""
        return self.pflat.data_ptr() + 4 * self.slots[name][0]

    def grad_ptr(self, name):
        return self.gflat.data_ptr() + 4 * self.slots[name][0]

    def valid(self):
        """False once the module's tensors were replaced (model.to(), .half(), ...): the engine then re-flattens."""
        return self._probe is None or self._probe.data_ptr() == self._probe_ptr

    def attach_grads(self, model_params, clone=False):
        """`.grad` of every trainable parameter := its view of the flat gradient buffer."""
        src = self.gflat.clone() if clone else self.gflat
        for n in self.names:
            o, k, shape = self.slots[n]
            model_params[n].grad = src[o:o + k].view(shape)
