"""tests/golden/make_golden_evalpost.py -- golden COCO-json rows from the UNMODIFIED reference's Evaler post-processing
(yolov6/core/evaler.py:324-384) on seeded synthetic NMS outputs (build container only).

    PYTHONPATH=tests/golden/refshim:/root/reference:. python tests/golden/make_golden_evalpost.py
"""
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "refshim"), "/root/reference", ROOT]
for name in ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval"):      # evaler.py imports them at module level
    mod = types.ModuleType(name)
    mod.COCO = mod.COCOeval = object
    sys.modules.setdefault(name, mod)

import torch  # noqa: E402

torch.cuda.is_available = lambda: False
from yolov6.core.evaler import Evaler  # noqa: E402

from oracle import evalpost as oe  # noqa: E402

if __name__ == "__main__":
    ev = Evaler.__new__(Evaler)          # only the stateless post-processing methods are used
    ev.is_coco = True
    ids = Evaler.coco80_to_coco91_class()
    store = {}
    for seed in (0, 1):
        outs, paths, shapes = oe.synthetic_batch(seed=seed)
        imgs = [torch.zeros(3, 640, 640)] * len(outs)
        rows = ev.convert_to_coco_format([o.clone() for o in outs], imgs, paths, shapes, ids)
        store[f"seed{seed}"] = rows
        print("seed", seed, len(rows), "rows")
    with open(os.path.join(HERE, "evalpost.json"), "w") as f:
        json.dump(store, f)
