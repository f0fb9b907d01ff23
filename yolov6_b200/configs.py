"""Built-in model configurations (the `model` dict of reference configs/yolov6{n,s,m,l6}.py) so that
tests, smoke() and bench.py run where /root/reference is not mounted, plus a normaliser that accepts
the reference's own mmcv-style Config object (yolov6/utils/config.py) for drop-in use."""
import copy

_P5_HEAD = dict(type="EffiDeHead", in_channels=[128, 256, 512], num_layers=3, begin_indices=24, anchors=3,
                anchors_init=[[10, 13, 19, 19, 33, 23], [30, 61, 59, 59, 59, 119], [116, 90, 185, 185, 373, 326]],   # fuse_ab head
                out_indices=[17, 20, 23], strides=[8, 16, 32], atss_warmup_epoch=0)

CONFIGS = {
    "yolov6n": dict(
        training_mode="repvgg", depth_multiple=0.33, width_multiple=0.25,
        backbone=dict(type="EfficientRep", num_repeats=[1, 6, 12, 18, 6], out_channels=[64, 128, 256, 512, 1024],
                      fuse_P2=True, cspsppf=True),
        neck=dict(type="RepBiFPANNeck", num_repeats=[12, 12, 12, 12], out_channels=[256, 128, 128, 256, 256, 512]),
        head=dict(_P5_HEAD, iou_type="siou", use_dfl=False, reg_max=0)),
    "yolov6s": dict(
        training_mode="repvgg", depth_multiple=0.33, width_multiple=0.50,
        backbone=dict(type="EfficientRep", num_repeats=[1, 6, 12, 18, 6], out_channels=[64, 128, 256, 512, 1024],
                      fuse_P2=True, cspsppf=True),
        neck=dict(type="RepBiFPANNeck", num_repeats=[12, 12, 12, 12], out_channels=[256, 128, 128, 256, 256, 512]),
        head=dict(_P5_HEAD, iou_type="giou", use_dfl=False, reg_max=0)),
    "yolov6m": dict(
        training_mode="repvgg", depth_multiple=0.60, width_multiple=0.75,
        backbone=dict(type="CSPBepBackbone", num_repeats=[1, 6, 12, 18, 6], out_channels=[64, 128, 256, 512, 1024],
                      csp_e=2.0 / 3, fuse_P2=True),
        neck=dict(type="CSPRepBiFPANNeck", num_repeats=[12, 12, 12, 12], out_channels=[256, 128, 128, 256, 256, 512],
                  csp_e=2.0 / 3),
        head=dict(_P5_HEAD, iou_type="giou", use_dfl=True, reg_max=16)),
    "yolov6l6": dict(
        training_mode="conv_silu", depth_multiple=1.0, width_multiple=1.0,
        backbone=dict(type="CSPBepBackbone_P6", num_repeats=[1, 6, 12, 18, 6, 6],
                      out_channels=[64, 128, 256, 512, 768, 1024], csp_e=0.5, fuse_P2=True),
        neck=dict(type="CSPRepBiFPANNeck_P6", num_repeats=[12, 12, 12, 12, 12, 12],
                  out_channels=[512, 256, 128, 256, 512, 1024], csp_e=0.5),
        head=dict(type="EffiDeHead", in_channels=[128, 256, 512, 1024], num_layers=4, anchors=1,
                  strides=[8, 16, 32, 64], atss_warmup_epoch=4, iou_type="giou", use_dfl=True, reg_max=16)),
}


def get_config(name):
    return copy.deepcopy(CONFIGS[name])


def normalize(cfg):
    """Accept a built-in name, one of the dicts above, or the reference's Config (attribute access,
    `.model.{depth_multiple,width_multiple,backbone,neck,head}`, `.training_mode`) -> plain dict."""
    if isinstance(cfg, str):
        return get_config(cfg)
    if isinstance(cfg, dict) and "depth_multiple" in cfg:
        return copy.deepcopy(cfg)
    model = cfg["model"] if isinstance(cfg, dict) else cfg.model
    get = (lambda o, k, d=None: o.get(k, d)) if hasattr(model, "get") else (lambda o, k, d=None: getattr(o, k, d))
    mode = (cfg.get("training_mode") if isinstance(cfg, dict) else getattr(cfg, "training_mode", None)) or "repvgg"
    out = dict(training_mode=mode, depth_multiple=get(model, "depth_multiple"), width_multiple=get(model, "width_multiple"))
    for part in ("backbone", "neck", "head"):
        out[part] = {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in dict(get(model, part)).items()}
    return out
