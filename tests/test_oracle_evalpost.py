"""Pins oracle/evalpost.py against the reference's Evaler.convert_to_coco_format rows (tests/golden/evalpost.json)."""
from conftest import golden_json
from oracle import evalpost as oe

IDS = [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 27, 28, 31, 32, 33, 34, 35, 36, 37, 38, 39, 40,
       41, 42, 43, 44, 46, 47, 48, 49, 50, 51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63, 64, 65, 67, 70, 72, 73, 74, 75, 76, 77, 78,
       79, 80, 81, 82, 84, 85, 86, 87, 88, 89, 90]     # coco80_to_coco91_class (evaler.py)


def test_oracle_rows_equal_reference_rows():
    g = golden_json("evalpost.json")
    for seed in (0, 1):
        outs, paths, shapes = oe.synthetic_batch(seed=seed)
        rows = oe.convert_to_coco_format(outs, paths, shapes, IDS)
        assert rows == g[f"seed{seed}"]
