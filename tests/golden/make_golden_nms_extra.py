"""tests/golden/make_golden_nms_extra.py -- extra NMS golden vectors from the UNMODIFIED reference for regimes the main
cases do not reach; consumed by the CPU oracle test only (tests/test_oracle_nms.py):

  * more than max_nms = 30000 candidates in one image (nms.py:55,90-91): only the 30000 best by confidence enter
    torchvision.ops.nms.  Scores are continuous random numbers, so the reference's unstable argsort has no ties to break.
    (The CUDA path raises above 65536 multi-label candidates per image instead -- DESIGN.md section 5.)
  * multi_label + class filter + agnostic together (nms.py:75-77,86-87,94).

    PYTHONPATH=tests/golden/refshim:/root/reference:. python tests/golden/make_golden_nms_extra.py
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]
torch.cuda.is_available = lambda: False

from yolov6.utils.nms import non_max_suppression  # noqa: E402

from oracle import fabricate as fab  # noqa: E402

CASES = [  # (B, A, nc, seed, kwargs)
    (1, 4000, 80, 11, dict(conf_thres=0.2, iou_thres=0.65, multi_label=True, max_det=300)),      # ~105 k candidates -> 30000
    (2, 3000, 20, 12, dict(conf_thres=0.05, iou_thres=0.5, multi_label=True, agnostic=True, classes=[1, 2, 7, 19], max_det=100)),
]

if __name__ == "__main__":
    store = {}
    for i, (B, A, nc, seed, kw) in enumerate(CASES):
        p = fab.synthetic_predictions(B, A, nc, seed)
        ncand = int(((p[..., 5:] * p[..., 4:5]) > kw["conf_thres"]).sum())
        out = non_max_suppression(p.clone(), **kw)
        store[f"c{i}_checksum"] = np.float64(fab.checksum(p))
        store[f"c{i}_counts"] = np.array([o.shape[0] for o in out], dtype=np.int64)
        store[f"c{i}_rows"] = torch.cat(out).numpy() if sum(o.shape[0] for o in out) else np.zeros((0, 6), np.float32)
        print("case", i, "candidates", ncand, "kept", [o.shape[0] for o in out])
    np.savez_compressed(os.path.join(HERE, "nms_extra.npz"), **store)
    with open(os.path.join(HERE, "nms_extra_cases.json"), "w") as f:
        json.dump([[B, A, nc, seed, kw] for B, A, nc, seed, kw in CASES], f)
