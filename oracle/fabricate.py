"""oracle/fabricate.py -- deterministic synthetic checkpoints (TEST INFRASTRUCTURE ONLY).

There is no network for released weights, and random-init YOLOv6 heads emit ~0.01 scores
(Detect.initialize_biases, reference effidehead.py:49-65) so NMS would see nothing.  This builds a
train-form `state_dict` from nothing but the (key, shape) list of a model: every tensor is drawn
from a generator seeded by crc32(key) ^ seed, so the golden-vector generator (which runs the real
reference) and the tests (which run the oracle and the CUDA path) obtain bit-identical weights
without shipping them.  BatchNorm statistics and affine terms are randomised so that BN folding and
RepVGG re-parameterisation are non-trivial.
"""
import zlib

import torch


def _gen(key, seed):
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def fabricate_tensor(key, shape, seed=0):
    g = _gen(key, seed)
    shape = tuple(shape)
    if key.endswith("num_batches_tracked"):
        return torch.zeros((), dtype=torch.int64)
    if key.endswith("running_mean"):
        return torch.randn(shape, generator=g) * 0.1
    if key.endswith("running_var"):
        return torch.rand(shape, generator=g) + 0.5
    if key.endswith("alpha"):
        return torch.rand(shape, generator=g) + 0.5
    if key == "detect.proj":
        return torch.linspace(0, shape[0] - 1, shape[0])
    if key == "detect.proj_conv.weight":
        return torch.linspace(0, shape[1] - 1, shape[1]).view(shape)
    if ".cls_preds." in key:
        return torch.randn(shape, generator=g) * 0.05 if key.endswith("weight") else torch.full(shape, -2.0)
    if ".reg_preds." in key:
        return torch.randn(shape, generator=g) * 0.05 if key.endswith("weight") else torch.full(shape, 1.0)
    if len(shape) == 4:  # conv / conv-transpose weights: He-style scale keeps activations O(1)
        fan_in = shape[1] * shape[2] * shape[3]
        return torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    if key.endswith("weight"):  # BN gamma
        return torch.rand(shape, generator=g) + 0.5
    if key.endswith("bias"):    # BN beta / conv-transpose bias
        return torch.randn(shape, generator=g) * 0.1
    raise KeyError(f"fabricate: unrecognised parameter {key} {shape}")


def fabricate_state_dict(keys_shapes, seed=0):
    """keys_shapes: iterable of (key, shape).  Returns {key: fp32 tensor} (int64 for counters)."""
    return {k: fabricate_tensor(k, s, seed) for k, s in keys_shapes}


def synthetic_images(batch, height, width, seed=0):
    """Seeded uniform [0,1) images, NCHW fp32."""
    return torch.rand(batch, 3, height, width, generator=torch.Generator().manual_seed(1000 + seed))
