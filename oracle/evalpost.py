"""oracle/evalpost.py -- CPU restatement of the reference's evaluation post-processing (TEST INFRASTRUCTURE ONLY).

Follows Evaler.scale_coords / box_convert / convert_to_coco_format (reference yolov6/core/evaler.py:324-384) step by
step in torch fp32 on CPU; pinned against the live reference by tests/golden/make_golden_evalpost.py.
"""
from pathlib import Path

import torch


def scale_coords(coords, img0_shape, ratio_pad):
    """evaler.py:333-359 (tensor branch); coords [n,4] xyxy fp32, modified in place like the reference."""
    gain, pad = ratio_pad
    coords[:, [0, 2]] -= pad[0]
    coords[:, [0, 2]] /= gain[1]
    coords[:, [1, 3]] -= pad[1]
    coords[:, [1, 3]] /= gain[0]
    coords[:, 0].clamp_(0, img0_shape[1])
    coords[:, 1].clamp_(0, img0_shape[0])
    coords[:, 2].clamp_(0, img0_shape[1])
    coords[:, 3].clamp_(0, img0_shape[0])
    return coords


def box_convert(x):
    """evaler.py:324-331."""
    y = x.clone()
    y[:, 0] = (x[:, 0] + x[:, 2]) / 2
    y[:, 1] = (x[:, 1] + x[:, 3]) / 2
    y[:, 2] = x[:, 2] - x[:, 0]
    y[:, 3] = x[:, 3] - x[:, 1]
    return y


def convert_to_coco_format(outputs, paths, shapes, ids, is_coco=True):
    """evaler.py:361-384; outputs: list of [k,6] fp32 tensors (xyxy, conf, cls)."""
    results = []
    for i, pred in enumerate(outputs):
        if len(pred) == 0:
            continue
        pred = pred.clone().float()
        shape = shapes[i][0]
        scale_coords(pred[:, :4], shape, shapes[i][1])
        stem = Path(paths[i]).stem
        image_id = int(stem) if is_coco else stem
        bboxes = box_convert(pred[:, 0:4])
        bboxes[:, :2] -= bboxes[:, 2:] / 2
        cls, scores = pred[:, 5], pred[:, 4]
        for ind in range(pred.shape[0]):
            results.append({"image_id": image_id, "category_id": ids[int(cls[ind])],
                            "bbox": [round(x, 3) for x in bboxes[ind].tolist()], "score": round(scores[ind].item(), 5)})
    return results


def synthetic_batch(B=4, max_det=50, seed=0, img=640):
    """Seeded NMS-like outputs + letterbox metadata: list of [k,6] tensors, paths, shapes."""
    g = torch.Generator().manual_seed(4000 + seed)
    outs, paths, shapes = [], [], []
    for b in range(B):
        k = int(torch.randint(0, max_det + 1, (1,), generator=g)) if b != 1 else 0      # image 1: no detections
        xy = torch.rand(k, 2, generator=g) * (img - 80) - 20                           # some boxes start outside the image
        wh = torch.rand(k, 2, generator=g) * 200 + 2
        conf = torch.rand(k, 1, generator=g)
        cls = torch.randint(0, 80, (k, 1), generator=g).float()
        outs.append(torch.cat([xy, xy + wh, conf, cls], 1).float())
        h0, w0 = int(torch.randint(300, 900, (1,), generator=g)), int(torch.randint(300, 900, (1,), generator=g))
        r = min(img / h0, img / w0)
        pad = ((img - w0 * r) / 2, (img - h0 * r) / 2)
        shapes.append(((h0, w0), ((r * 1.0, r * 1.0), pad)))
        paths.append(f"/data/coco/images/val2017/{100000 + 37 * b:012d}.jpg")
    return outs, paths, shapes
