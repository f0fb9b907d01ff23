"""Synthetic checkpoints for benchmarking without network access.

`randomize_(model, seed)` fills a freshly built model with seeded weights whose activations stay O(1)
through ~70 layers and whose head emits a realistic spread of scores (a random-init YOLOv6 head emits
~0.01 everywhere -- Detect.initialize_biases, reference effidehead.py:49-65 -- so NMS would have nothing
to do).  BatchNorm statistics are randomised so that folding is non-trivial.
"""
import torch


@torch.no_grad()
def randomize_(model, seed=0):
    g = torch.Generator().manual_seed(seed)
    for name, t in model.state_dict().items():
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked") or name.startswith("detect.proj"):
            continue
        if name.endswith("running_mean"):
            v = torch.randn(shape, generator=g) * 0.1
        elif name.endswith("running_var"):
            v = torch.rand(shape, generator=g) + 0.5
        elif name.endswith("alpha"):
            v = torch.rand(shape, generator=g) + 0.5
        elif ".cls_preds." in name:
            # wide logits far below zero: on the uniform-noise images of the benchmark
            # about 6000 (anchor, class) pairs per 640x640 image pass the eval threshold 0.03 and a few dozen
            # pass 0.4 -- the NMS load of a trained detector (measured with the oracle)
            v = torch.randn(shape, generator=g) * (113.0 / shape[1] ** 0.5) if name.endswith("weight") else torch.full(shape, -23.8)
        elif ".reg_preds." in name:
            v = torch.randn(shape, generator=g) * (2.0 / shape[1] ** 0.5) if name.endswith("weight") else torch.full(shape, 1.0)
        elif len(shape) == 4:
            v = torch.randn(shape, generator=g) * (1.0 / (shape[1] * shape[2] * shape[3])) ** 0.5
        elif name.endswith("weight"):
            v = torch.rand(shape, generator=g) * 0.4 + 0.35
        else:
            v = torch.randn(shape, generator=g) * 0.1
        t.copy_(v.to(t.device, t.dtype))
    return model


def synthetic_targets(batch, seed=1, mean_per_image=7.3, num_classes=80, max_per_image=60):
    """COCO-shaped synthetic labels: per image n ~ clip(Poisson(7.3), 1, 60); cls ~ U{0..nc-1};
    w, h ~ U(.02, .6); centres uniform such that the box stays inside.  Rows (img, cls, cx, cy, w, h), fp32, normalised --
    the layout of the reference's collate_fn (yolov6/data/datasets.py:299-304)."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(batch):
        n = int(torch.poisson(torch.tensor([mean_per_image]), generator=g).clamp(1, max_per_image).item())
        cls = torch.randint(0, num_classes, (n, 1), generator=g).float()
        wh = torch.rand(n, 2, generator=g) * 0.58 + 0.02
        cxy = wh / 2 + torch.rand(n, 2, generator=g) * (1 - wh)
        rows.append(torch.cat([torch.full((n, 1), float(b)), cls, cxy, wh], 1))
    return torch.cat(rows).float()
