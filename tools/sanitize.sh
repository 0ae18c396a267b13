#!/bin/bash
# Run on the GPU box (gpurun): compute-sanitizer over the kernel-level parity tests (SURVEY 7.4 tier T7).
#   memcheck  : out-of-bounds / misaligned global, shared and local accesses, leaks of device allocations
#   racecheck : shared-memory hazards (the kernels with cp.async, mbarriers, TMEM and per-warp queues)
#   synccheck : illegal barrier / mbarrier / syncwarp use
# Logs: gpurun_out/<tag>_sanitizer_<tool>.log (copy the summaries to profiles/).  usage: tools/sanitize.sh [tag] [pytest -k expr]
cd "$(dirname "$0")/.."
TAG=${1:-r2}
SEL=${2:-"ingest or pyramid or gftt or lk or fit or warp_golden or featprep_noise or hamming or l2_match or size_filter"}
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check full"
  timeout 1500 $SAN --tool $tool $extra --error-exitcode 9 --print-limit 20 --log-file gpurun_out/${TAG}_sanitizer_${tool}.raw \
      python -m pytest tests/test_gpu_kernels.py tests/test_gpu_orb.py -x -q -m gpu -k "$SEL" -p no:cacheprovider \
      > gpurun_out/${TAG}_sanitizer_${tool}.pytest 2>&1
  rc=$?
  { echo "tool=$tool exit=$rc (9 = sanitizer errors)  tests: $(tail -n 1 gpurun_out/${TAG}_sanitizer_${tool}.pytest)";
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY|=========  *(Invalid|Race|Hazard|Barrier|Leaked)" gpurun_out/${TAG}_sanitizer_${tool}.raw | sort | uniq -c | head -40; } \
      > gpurun_out/${TAG}_sanitizer_${tool}.log
  cat gpurun_out/${TAG}_sanitizer_${tool}.log
done
