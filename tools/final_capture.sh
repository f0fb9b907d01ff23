#!/bin/bash
# Round-end evidence run (one gpurun call): GPU tests, bench, per-layer conv table (+ cuDNN beside it), in-situ timeline, ncu launch
# list, smoke.  Everything lands in gpurun_out/ and is copied to profiles/ afterwards.  Ordered by importance: a call that runs out
# of GPU budget still leaves the first files.  FULL=1 adds the reference arm and one `ncu --set full` capture of the 256@40 conv.
set -u
R=${1:-r01}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${R}_pytest_gpu.log | cut -c1-200
timeout 300 python bench.py > gpurun_out/${R}_bench_line.json 2> gpurun_out/${R}_bench.err; cut -c1-400 gpurun_out/${R}_bench_line.json
timeout 200 python tools/profile_layers.py > gpurun_out/${R}_conv_layers_yolov6s.md 2> /dev/null; sed -n 3,3p gpurun_out/${R}_conv_layers_yolov6s.md
timeout 200 python tools/conv_vs_cudnn.py > gpurun_out/${R}_conv_vs_cudnn.md 2> /dev/null; tail -1 gpurun_out/${R}_conv_vs_cudnn.md | cut -c1-200
timeout 200 python tools/trace_step.py > gpurun_out/${R}_timeline_yolov6s.md 2> /dev/null; tail -2 gpurun_out/${R}_timeline_yolov6s.md
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  --kernel-name-base demangled -k regex:yv6:: -c ${NCU_LAUNCHES:-300} --csv --log-file gpurun_out/${R}_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
wc -l gpurun_out/${R}_launches_bench.csv
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | cut -c1-200
if [ "${FULL:-0}" = "1" ]; then
  timeout 600 python bench.py --impl reference > gpurun_out/${R}_bench_reference_line.json 2>> gpurun_out/${R}_bench.err; cut -c1-300 gpurun_out/${R}_bench_reference_line.json
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_igemm -s 2 -c 1 \
    -o gpurun_out/${R}_conv_256at40_full -f python tools/gpu_conv_check.py --profile "s1_256@40" relu > /dev/null 2>&1
fi
ls -la gpurun_out | tail -12
