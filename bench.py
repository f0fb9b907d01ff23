"""bench.py -- images/sec of the YOLOv6-S 640x640 bs32 inference hot path on N x B200 (BASELINE.json).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]

One "step" = one batch through the whole hot path: stem -> backbone -> neck -> head -> decode
(sm_100a kernels via the C ABI) -> batched NMS (eval settings conf 0.03 / iou 0.65 / multi_label,
the settings of the reference's Evaler, core/evaler.py:118-134).  Per rank the batch is 32 images
(weak scaling: images are sharded across GPUs, no collective on this path -- SURVEY.md 8e).

JSON keys (see the task contract): value = device-resident throughput (CUDA events, max over ranks);
e2e = same metric through the public API from pinned HOST uint8 images incl. H2D and the D2H of the
detections; roofline = algorithmic conv FLOPs / measured conv-kernel time vs the measured bf16 peak;
cpu_baseline = the oracle (CPU restatement of the reference path) on a bounded sample of the workload.
`--impl reference` times that CPU path alone with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec YOLOv6-S 640 bs32"
GFLOP_PER_IMG = {"yolov6n": 11.316, "yolov6s": 44.967, "yolov6m": 85.087, "yolov6l6": 665.834}
NMS_KW = dict(conf_thres=0.03, iou_thres=0.65, multi_label=True, max_det=300)


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def load_keys(model):
    with open(os.path.join(ROOT, "tests", "golden", f"keys_{model}.json")) as f:
        return [(k, tuple(s)) for k, s in json.load(f)]


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows if len(r) > 2 + i)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def host_cores():
    """Host threads this process can really use: CPU affinity, capped by the cgroup CPU quota (a container
    that sees 128 logical CPUs but is throttled to a few would otherwise be timed oversubscribed)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:                       # cgroup v2
            quota, period = f.read().split()
        if quota != "max":
            n = max(1, min(n, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:                                                            # cgroup v1
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
                quota = int(f.read())
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                period = int(f.read())
            if quota > 0:
                n = max(1, min(n, quota // period))
        except (OSError, ValueError):
            pass
    return n


def pick_threads(step_one_image):
    """The thread count (<= host_cores()) at which the CPU path runs fastest on a 1-image probe -- the CPU arm
    gets its best configuration, not an oversubscribed one."""
    best, best_t = None, float("inf")
    cores = host_cores()
    for n in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True):
        torch.set_num_threads(n)
        step_one_image()
        t0 = time.perf_counter()
        step_one_image()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def cpu_reference_step(sd, cfg, images, nms_kw):
    """The reference path restated on CPU (oracle): eval forward + NMS on `images` (NCHW fp32)."""
    from oracle import model as om
    from oracle import nms as onms
    with torch.no_grad():
        pred = om.forward(sd, cfg, images)
    return onms.non_max_suppression(pred.numpy(), **nms_kw)


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port; /root/reference is
    not on the GPU box) on the host cores; rank 0 only."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    from oracle import fabricate as fab
    from oracle import model as om
    from yolov6_b200.model import build_model          # only as the container of the seeded synthetic checkpoint
    from yolov6_b200.synth import randomize_            # (same weights as the GPU arm); nothing of it is timed
    sd = {k: v.detach() for k, v in randomize_(build_model(args.model, 80, torch.device("cpu")), seed=0).state_dict().items()}
    cfg = om.CONFIGS[args.model]
    sample = args.ref_batch
    x = fab.synthetic_images(sample, args.size, args.size, seed=0)
    cores = pick_threads(lambda: cpu_reference_step(sd, cfg, x[:1], NMS_KW))
    for _ in range(args.warmup_ref):
        cpu_reference_step(sd, cfg, x, NMS_KW)
    t0 = time.perf_counter()
    for _ in range(args.steps_ref):
        cpu_reference_step(sd, cfg, x, NMS_KW)
    dt = (time.perf_counter() - t0) / args.steps_ref
    val = sample / dt
    line = {"metric": METRIC, "value": val, "unit": "images/s", "n_gpus": world, "steps": args.steps_ref, "warmup": args.warmup_ref,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{args.model} {args.size}x{args.size} inference + NMS (CPU oracle port of the reference path)",
                       "sample": f"batch {sample} per step", "nms": NMS_KW},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps_ref} steps x batch {sample} at {args.size}x{args.size}"},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="yolov6s")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--steps-ref", type=int, default=3)
    ap.add_argument("--warmup-ref", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, local_rank, world = dist_env()
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from yolov6_b200.model import build_model
    from yolov6_b200.nms import nms_batched
    from yolov6_b200.synth import randomize_

    model = randomize_(build_model(args.model, 80, dev), seed=0)   # seeded synthetic checkpoint
    model.eval().set_precision(args.precision)
    eng = model.engine()
    B, S = args.batch, args.size
    g = torch.Generator().manual_seed(1 + rank)   # per-rank data like tools/train.py:104 seeds per rank
    host_u8 = [(torch.rand(B, 3, S, S, generator=g) * 255).to(torch.uint8).pin_memory() for _ in range(2)]
    dev_f32 = [h.to(dev).float() / 255 for h in host_u8]   # 157 MB each > 126 MB L2

    from yolov6_b200.pipeline import DetectPipeline
    use_graph = not args.no_graph
    if use_graph:
        # steady state = one CUDA-graph launch per batch (yolov6_b200/pipeline.py); two pipelines with
        # separate static buffers alternate so that consecutive steps never reuse a cached input
        pipes_dev = [DetectPipeline(model, B, S, S, host_input=False, **NMS_KW) for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)   # H2D of batch i+1 overlaps the kernels of batch i
        pipes_e2e = [DetectPipeline(model, B, S, S, host_input=True, overlap_h2d=True, copy_stream=copy_stream, **NMS_KW)
                     for _ in range(2)]
        for i in range(2):
            pipes_dev[i].x_dev.copy_(dev_f32[i])
            pipes_e2e[i].x_host.copy_(host_u8[i])

        def step_device(i):
            pipes_dev[i & 1].launch()

        def step_e2e(i):
            pipes_e2e[i & 1].launch()          # H2D (39 MB u8, copy stream) -> graph: kernels -> D2H detections
    else:
        def step_device(i):
            pred = eng.forward(dev_f32[i & 1])
            return nms_batched(pred, **NMS_KW)

        def step_e2e(i):
            x = host_u8[i & 1].to(dev, non_blocking=True)
            pred = eng.forward(x)
            out, count, _, _ = nms_batched(pred, **NMS_KW)
            return out.cpu(), count.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    with torch.no_grad():
        sampler = ClockSampler(local_rank)
        sampler.start()
        ms_dev = timed(step_device, args.steps, max(args.warmup, 3))
        ms_e2e = timed(step_e2e, args.steps, max(args.warmup, 3))
        sampler.stop_flag = True                # sampled across both timed regions
        # roofline of the dominant kernel (conv_igemm): CUDA events around every conv launch of 5 steps
        conv_ms, conv_flop, n_conv = eng.profile_convs(dev_f32[0], steps=5)
    sampler.join(timeout=2)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        with open(pk_path) as f:
            peaks = json.load(f)
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    achieved_tf = conv_flop / (conv_ms * 1e-3) / 1e12
    nms_launches = 3              # nms_select, nms_sort, nms_greedy
    traffic = None                # DRAM bytes per conv launch from the committed ncu capture (profiles/)
    tpath = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": world * B / (ms_dev * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (fp32-equivalent)", "data": "synthetic",
        "config": {"workload": f"{args.model} {S}x{S} bs{B}/GPU inference: forward + decode + batched NMS",
                   "nms": NMS_KW, "weights": "seeded random (yolov6_b200/synth.py)", "parallelism": f"dp{world} image-sharded, no collective",
                   "l2": "inputs (157 MB fp32 per batch, two alternating buffers) exceed the 126 MB L2",
                   "launch": "one CUDA graph per batch (DetectPipeline)" if use_graph else "eager ctypes launches",
                   "e2e_pipeline": "two alternating pipelines; the pinned-host -> device copy of a batch runs on a copy stream and "
                                   "overlaps the kernels of the previous batch; detections are copied back inside the graph"},
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": "images/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": B * 3 * S * S, "d2h_bytes_per_step": B * NMS_KW["max_det"] * 6 * 4 + B * 4},
        "gpu_launches": (eng.launch_count(B, S, S) + nms_launches) * args.steps,
        "roofline": {"bound": "tensor", "kernel": "yv6::conv_igemm_kernel", "achieved": achieved_tf, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                     "traffic_note": "bytes per conv_igemm launch, ncu dram__bytes_read+write averaged over the step's launches "
                                     "(profiles/r01_conv_traffic.json); algorithmic activation bytes per launch = "
                                     f"{eng.conv_bytes_per_launch(B, S, S):.3e}",
                     "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s (B200_PROFILING.md)",
                     "launches_per_step": n_conv, "conv_ms_per_step": conv_ms, "algorithmic_gflop_per_step": conv_flop / 1e9,
                     "model_gflop_per_img": GFLOP_PER_IMG.get(args.model)},
        "clocks": sampler.summary(),
    }
    if not args.no_cpu_baseline and world == 1:
        from oracle import model as om        # CPU-baseline leg: the checker, timed on the host cores
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        xs = torch.rand(args.ref_batch, 3, S, S, generator=torch.Generator().manual_seed(0))
        cores = pick_threads(lambda: cpu_reference_step(sd, om.CONFIGS[args.model], xs[:1], NMS_KW))
        t0 = time.perf_counter()
        for _ in range(args.steps_ref):
            cpu_reference_step(sd, om.CONFIGS[args.model], xs, NMS_KW)
        dt = (time.perf_counter() - t0) / args.steps_ref
        line["cpu_baseline"] = {"value": args.ref_batch / dt, "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"{args.steps_ref} steps x batch {args.ref_batch} of the same workload (fp32 oracle: forward + NMS)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
