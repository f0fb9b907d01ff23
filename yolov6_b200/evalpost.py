"""Evaluation post-processing on top of the batched NMS output.

The reference's `Evaler.convert_to_coco_format` (yolov6/core/evaler.py:361-384) walks the detections of every image
in Python: `scale_coords` back to the original image (:333-359), `box_convert` to centre form (:324-331), shift to the
top-left corner, and one `round()` + dict per detection.  Here the geometry of ALL images is one kernel launch over the
fixed-size NMS output (`yv6_eval_boxes`), one device->host copy follows, and the json rows are built from a flat list.
`to_end2end` exposes the same NMS output in the tensor layout of the reference's ONNX / TensorRT `End2End` wrappers
(yolov6/models/end2end.py: num_dets, det_boxes, det_scores, det_classes).
"""
import ctypes as C
from pathlib import Path

import torch

from . import _lib


def image_meta(shapes, device):
    """shapes: the dataloader's per-image `(h0, w0), ((h_ratio, w_ratio), (pad_w, pad_h))` (yolov6/data/datasets.py) ->
    [B,6] fp32 (gain_h, gain_w, pad_x, pad_y, h0, w0)."""
    rows = []
    for shape, (gain, pad) in shapes:
        rows.append([float(gain[0]), float(gain[1]), float(pad[0]), float(pad[1]), float(shape[0]), float(shape[1])])
    return torch.tensor(rows, dtype=torch.float32).to(device)


def eval_boxes(out, count, meta, stream=None):
    """out [B,max_det,6] xyxy/conf/cls + count [B] (from nms_batched) -> [B,max_det,6] (x_tl, y_tl, w, h, conf, cls) in the
    original images' pixels; stays on the device, no host synchronisation."""
    if out.device.type != "cuda":
        raise RuntimeError("yolov6_b200.evalpost runs on CUDA tensors only (no CPU fallback)")
    B, max_det, _ = out.shape
    res = torch.empty_like(out)
    _lib.check(_lib.lib().yv6_eval_boxes(_lib.handle(out.device.index or 0), C.c_void_p(out.data_ptr()), C.c_void_p(count.data_ptr()),
                                         C.c_void_p(meta.data_ptr()), B, max_det, C.c_void_p(res.data_ptr()), _lib.stream_ptr(stream)))
    return res


def convert_to_coco_format(out, count, paths, shapes, ids, is_coco=True):
    """Drop-in result of Evaler.convert_to_coco_format for a batch, from the batched NMS tensors."""
    meta = image_meta(shapes, out.device)
    boxes = eval_boxes(out.contiguous().float(), count, meta)
    host = torch.cat([boxes.reshape(boxes.shape[0], -1), count.view(-1, 1).float()], 1).cpu()     # the one D2H copy
    counts = host[:, -1].long().tolist()
    rows = host[:, :-1].reshape(boxes.shape).tolist()
    results = []
    for i, n in enumerate(counts):
        if n == 0:
            continue
        stem = Path(paths[i]).stem
        image_id = int(stem) if is_coco else stem
        for r in rows[i][:n]:
            results.append({"image_id": image_id, "category_id": ids[int(r[5])],
                            "bbox": [round(v, 3) for v in r[:4]], "score": round(r[4], 5)})
    return results


def to_end2end(out, count):
    """(num_dets [B,1] int32, det_boxes [B,max_det,4], det_scores [B,max_det], det_classes [B,max_det] int32): the output
    signature of the reference's End2End / TRT EfficientNMS wrappers (yolov6/models/end2end.py)."""
    return count.view(-1, 1).to(torch.int32), out[..., :4], out[..., 4], out[..., 5].to(torch.int32)
