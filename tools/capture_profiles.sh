#!/bin/bash
# Run on the GPU box (gpurun): the ncu evidence committed under profiles/ (round tag = $1, default r1).
#   1. launch list of the bench command (gpu__time_duration.sum, clocks untouched)
#   2. one `--set full` capture of each hot kernel of the LK + warp path (one launch = one batch of frames)
#   3. launch list of streaming calls (per-frame kernels)
cd "$(dirname "$0")/.."
TAG=${1:-r2}
ARGS="--steps 2 --warmup 1 --frames-per-gpu 256 --batch 256 --no-cpu-baseline --no-e2e --no-mode-probes --no-c5-probe --no-parity"
python bench.py $ARGS > gpurun_out/plain_${TAG}.log 2>&1 || { echo "bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_256f.csv \
    python bench.py $ARGS > gpurun_out/ncu_list_${TAG}.log 2>&1
ncu --set full --import-source on --clock-control none \
    -k regex:'ingest_kernel|pyrdown_kernel|lkprep_kernel|eig_kernel|topk_greedy_kernel|lk_chain_kernel|fit_kernel|smooth_kernel|warp_tile_kernel|acc_chunk' \
    -c 16 -o gpurun_out/${TAG}_full -f python bench.py $ARGS > gpurun_out/ncu_full_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_streaming.csv \
    python tools/stream_probe.py 56 > gpurun_out/stream_probe_${TAG}.log 2>&1
ls -la gpurun_out/${TAG}_*
