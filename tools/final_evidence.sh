#!/bin/bash
# The single-GPU evidence of a build, in one gpurun call: GPU test suite, smoke(), the default bench line, the reference arm,
# BASELINE config 5 with 100 000 4K frames, and an ncu capture of the simulator kernel.  Files land in gpurun_out/${TAG}_*.
# usage (repo root): gpurun --timeout 1500 -- 'bash tools/final_evidence.sh r2h'
cd "$(dirname "$0")/.."
TAG=${1:-r2h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_gputests.log 2>&1; tail -3 gpurun_out/${TAG}_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err; grep -c '^{' gpurun_out/${TAG}_bench.log
python bench.py --impl reference > gpurun_out/${TAG}_ref.log 2> gpurun_out/${TAG}_ref.err; grep -c '^{' gpurun_out/${TAG}_ref.log
python bench.py --workload c5 --c5-frames 100000 --steps 1 --warmup 1 > gpurun_out/${TAG}_c5_n1.log 2> gpurun_out/${TAG}_c5_n1.err; grep -c '^{' gpurun_out/${TAG}_c5_n1.log
ncu --set full --import-source on --clock-control none -k regex:'render_kernel' -c 2 -o gpurun_out/${TAG}_render -f \
  python bench.py --workload c5 --c5-frames 256 --steps 1 --warmup 0 > gpurun_out/${TAG}_ncu_render.log 2>&1
ls -la gpurun_out/${TAG}_*
