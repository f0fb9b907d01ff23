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
    if ".cls_preds." in key:  # logits with std ~1.5 around -2: scores spread over (0.01, 0.9)
        return torch.randn(shape, generator=g) * (6.0 / shape[1] ** 0.5) if key.endswith("weight") else torch.full(shape, -2.0)
    if ".reg_preds." in key:
        return torch.randn(shape, generator=g) * (2.0 / shape[1] ** 0.5) if key.endswith("weight") else torch.full(shape, 1.0)
    if len(shape) == 4:  # conv / conv-transpose weights: unit-gain scale, activations stay O(0.1..1)
        fan_in = shape[1] * shape[2] * shape[3]
        return torch.randn(shape, generator=g) * (1.0 / fan_in) ** 0.5
    if key.endswith("weight"):  # BN gamma: ~1/sqrt(3) so the three RepVGG branches do not blow up
        return torch.rand(shape, generator=g) * 0.4 + 0.35
    if key.endswith("bias"):    # BN beta / conv-transpose bias
        return torch.randn(shape, generator=g) * 0.1
    raise KeyError(f"fabricate: unrecognised parameter {key} {shape}")


def fabricate_state_dict(keys_shapes, seed=0):
    """keys_shapes: iterable of (key, shape).  Returns {key: fp32 tensor} (int64 for counters)."""
    return {k: fabricate_tensor(k, s, seed) for k, s in keys_shapes}


def synthetic_images(batch, height, width, seed=0):
    """Seeded uniform [0,1) images, NCHW fp32."""
    return torch.rand(batch, 3, height, width, generator=torch.Generator().manual_seed(1000 + seed))


def synthetic_predictions(batch, anchors, num_classes, seed=0, clusters=12, img=640.0):
    """Seeded eval-form predictions [B, A, 5+nc] (xywh, obj=1, cls) with boxes clustered around a few
    centres so that NMS really suppresses; scores ~ U^4 so a realistic fraction passes conf."""
    g = torch.Generator().manual_seed(2000 + seed)
    centers = torch.rand(batch, clusters, 2, generator=g) * (img - 40) + 20
    which = torch.randint(0, clusters, (batch, anchors), generator=g)
    cxy = torch.gather(centers, 1, which[..., None].expand(batch, anchors, 2)) + torch.randn(batch, anchors, 2, generator=g) * 6
    wh = torch.rand(batch, anchors, 2, generator=g) * 80 + 10
    cls = torch.rand(batch, anchors, num_classes, generator=g) ** 4
    return torch.cat([cxy, wh, torch.ones(batch, anchors, 1), cls], -1).float()


def synthetic_predictions_sparse(batch, anchors, num_classes, seed=0, density=0.012, clusters=12, img=640.0):
    """Like `synthetic_predictions`, but only a fraction `density` of the (anchor, class) scores is above the Evaler's
    conf 0.03 (the rest sits at ~1e-3): ~6000 candidates per 8400-anchor image, the load of the benchmark workload."""
    g = torch.Generator().manual_seed(2500 + seed)
    centers = torch.rand(batch, clusters, 2, generator=g) * (img - 40) + 20
    which = torch.randint(0, clusters, (batch, anchors), generator=g)
    cxy = torch.gather(centers, 1, which[..., None].expand(batch, anchors, 2)) + torch.randn(batch, anchors, 2, generator=g) * 6
    wh = torch.rand(batch, anchors, 2, generator=g) * 80 + 10
    hot = torch.rand(batch, anchors, num_classes, generator=g) < density
    cls = torch.where(hot, torch.rand(batch, anchors, num_classes, generator=g) ** 2 * 0.9 + 0.03,
                      torch.rand(batch, anchors, num_classes, generator=g) * 2e-3)
    return torch.cat([cxy, wh, torch.ones(batch, anchors, 1), cls], -1).float()


def synthetic_head_outputs(batch, sizes, num_classes, reg_ch, seed=0):
    """Seeded train-form head outputs: cls [B,A,nc] post-sigmoid, reg [B,A,reg_ch] (ltrb distances in
    stride units when reg_ch == 4, DFL logits otherwise)."""
    g = torch.Generator().manual_seed(3000 + seed)
    A = sum(h * w for h, w in sizes)
    cls = torch.sigmoid(torch.randn(batch, A, num_classes, generator=g) * 1.5 - 2.0)
    if reg_ch == 4:
        reg = torch.rand(batch, A, 4, generator=g) * 4 + 0.5
    else:
        reg = torch.randn(batch, A, reg_ch, generator=g) * 0.8
    return cls.float(), reg.float()


def checksum(t):
    """Order-sensitive fingerprint of a tensor, used to detect RNG drift between the golden generator
    and the test host before comparing against stored outputs."""
    t = t.detach().double().flatten()
    w = torch.arange(1, t.numel() + 1, dtype=torch.float64)
    return float((t * (w % 9973 + 1)).sum())
