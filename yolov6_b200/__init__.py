"""yolov6_b200 -- B200-native (sm_100a) kernels behind the YOLOv6 hot-path API."""
__version__ = "0.1.0"
