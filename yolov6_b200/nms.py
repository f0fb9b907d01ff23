"""Drop-in `non_max_suppression` backed by the batched sm_100a NMS kernels (csrc/yv6_nms.cu).

Same signature, assertions and return type as the reference's yolov6/utils/nms.py:31-105
(`list` of B tensors [k,6] = xyxy, conf, cls on the input device; empty -> zeros((0,6))), but all
images are processed by three kernel launches (select, sort, greedy) instead of a Python loop around torchvision.ops.nms.
"""
import ctypes as C

import torch

from . import _lib

_ws_cache = {}


def _workspace(dev, nbytes):
    buf = _ws_cache.get(dev)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
        _ws_cache[dev] = buf
    return buf


def workspace_bytes(B, A, nc, multi_label):
    return int(_lib.lib().yv6_nms_workspace_bytes(B, A, nc, 1 if (multi_label and nc > 1) else 0))


def nms_batched(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                max_det=300, stream=None, workspace=None):
    """Raw batched form: returns (out [B,max_det,6], count [B] int32, src [B,max_det,2] int32 (anchor, class),
    overflow [1] int32), all on the device, no host synchronisation.  `workspace` (uint8 device tensor of at least
    `workspace_bytes(...)`) lets a caller that captures the call into a CUDA graph own the scratch memory; eager calls
    share one growable buffer per device."""
    if prediction.device.type != "cuda":
        raise RuntimeError("yolov6_b200.non_max_suppression runs on CUDA tensors only (no CPU fallback)")
    assert 0 <= conf_thres <= 1, f'conf_thresh must be in 0.0 to 1.0, however {conf_thres} is provided.'
    assert 0 <= iou_thres <= 1, f'iou_thres must be in 0.0 to 1.0, however {iou_thres} is provided.'
    pred = prediction.contiguous()
    if pred.dtype != torch.float32:
        pred = pred.float()
    B, A, no = pred.shape
    nc = no - 5
    dev = pred.device
    lib = _lib.lib()
    ml = 1 if (multi_label and nc > 1) else 0
    nbytes = lib.yv6_nms_workspace_bytes(B, A, nc, ml)
    if workspace is not None:
        if workspace.device != dev or workspace.numel() < nbytes:
            raise RuntimeError(f"nms workspace too small ({workspace.numel()} < {nbytes} bytes) or on the wrong device")
        ws = workspace
    else:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("nms_batched inside a CUDA-graph capture needs a caller-owned `workspace` (the shared "
                               "per-device buffer may be replaced later while the graph still points at it)")
        ws = _workspace(dev, nbytes)
    out = torch.zeros(B, max_det, 6, dtype=torch.float32, device=dev)
    count = torch.zeros(B, dtype=torch.int32, device=dev)
    src = torch.zeros(B, max_det, 2, dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    mask_ptr = C.c_void_p(0)
    mask = None
    if classes is not None:
        mask = torch.zeros(nc, dtype=torch.uint8)
        mask[[int(c) for c in classes if 0 <= int(c) < nc]] = 1
        mask = mask.to(dev)
        mask_ptr = C.c_void_p(mask.data_ptr())
    _lib.check(lib.yv6_nms_batched(_lib.handle(dev.index or 0), C.c_void_p(pred.data_ptr()), B, A, nc,
                                   float(conf_thres), float(iou_thres), int(bool(agnostic)), ml, mask_ptr,
                                   int(max_det), C.c_void_p(out.data_ptr()), C.c_void_p(count.data_ptr()),
                                   C.c_void_p(src.data_ptr()), C.c_void_p(overflow.data_ptr()),
                                   C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr(stream)))
    return out, count, src, overflow


def nms_batched_head(cls, reg, sizes, strides, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                     max_det=300, stream=None, workspace=None):
    """`nms_batched` on the head tensors (cls [B,A,nc] post-sigmoid, reg [B,A,R], level grid `sizes` / `strides`) without the
    intermediate `[B,A,5+nc]` prediction tensor: boxes are decoded only for the candidates.  Same outputs, bit for bit, as
    `nms_batched(decode(cls, reg))` -- the serving pipeline's path (pipeline.py)."""
    if cls.device.type != "cuda":
        raise RuntimeError("yolov6_b200.non_max_suppression runs on CUDA tensors only (no CPU fallback)")
    assert 0 <= conf_thres <= 1, f'conf_thresh must be in 0.0 to 1.0, however {conf_thres} is provided.'
    assert 0 <= iou_thres <= 1, f'iou_thres must be in 0.0 to 1.0, however {iou_thres} is provided.'
    B, A, nc = cls.shape
    R = reg.shape[2]
    dev = cls.device
    lib = _lib.lib()
    ml = 1 if (multi_label and nc > 1) else 0
    nbytes = lib.yv6_nms_workspace_bytes(B, A, nc, ml)
    if workspace is not None:
        if workspace.device != dev or workspace.numel() < nbytes:
            raise RuntimeError(f"nms workspace too small ({workspace.numel()} < {nbytes} bytes) or on the wrong device")
        ws = workspace
    else:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("nms_batched_head inside a CUDA-graph capture needs a caller-owned `workspace`")
        ws = _workspace(dev, nbytes)
    out = torch.zeros(B, max_det, 6, dtype=torch.float32, device=dev)
    count = torch.zeros(B, dtype=torch.int32, device=dev)
    src = torch.zeros(B, max_det, 2, dtype=torch.int32, device=dev)
    overflow = torch.zeros(1, dtype=torch.int32, device=dev)
    mask_ptr, mask = C.c_void_p(0), None
    if classes is not None:
        mask = torch.zeros(nc, dtype=torch.uint8)
        mask[[int(c) for c in classes if 0 <= int(c) < nc]] = 1
        mask = mask.to(dev)
        mask_ptr = C.c_void_p(mask.data_ptr())
    nl = len(sizes)
    assert sum(h * w for h, w in sizes) == A
    lh = (C.c_int32 * nl)(*[int(h) for h, _ in sizes])
    lw = (C.c_int32 * nl)(*[int(w) for _, w in sizes])
    ls = (C.c_float * nl)(*[float(s) for s in strides])
    _lib.check(lib.yv6_nms_batched_head(_lib.handle(dev.index or 0), C.c_void_p(cls.data_ptr()), C.c_void_p(reg.data_ptr()), B, nc, R, nl,
                                        lh, lw, ls, float(conf_thres), float(iou_thres), int(bool(agnostic)), ml, mask_ptr, int(max_det),
                                        C.c_void_p(out.data_ptr()), C.c_void_p(count.data_ptr()), C.c_void_p(src.data_ptr()),
                                        C.c_void_p(overflow.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), _lib.stream_ptr(stream)))
    return out, count, src, overflow


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, max_det=300):
    out, count, _, overflow = nms_batched(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det)
    host = torch.cat([count, overflow]).tolist()          # the one host sync: detection counts
    if host[-1]:
        # images with more than 65536 candidates are reduced to their 30000 best on the device (nms.py:90-91); this
        # remains only for > 65536 candidates whose scores agree to 8 significant bits (see csrc/yv6_nms.cu)
        raise RuntimeError("non_max_suppression: more than 65536 near-identical candidate scores in one image; raise conf_thres")
    return [out[i, :host[i]] for i in range(out.shape[0])]
