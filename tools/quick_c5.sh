#!/bin/bash
# Quick check on the GPU box: the offline-runner tests, then BASELINE config 5 (4K, GLOBAL_SMOOTHING, simulator source) at 2048 frames:
# phases of the fused single pass against the two-pass schedule (VSTAB_OFFLINE_FUSED=0).  usage: gpurun -- 'bash tools/quick_c5.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_offline.py -q -m gpu -x 2>&1 | tail -4
for fused in 1 0; do
  VSTAB_OFFLINE_FUSED=$fused python bench.py --workload c5 --c5-frames 2048 --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['c5']
        print('fused=$fused value %.0f frames/s' % d['value'], 'phases', {k: round(v, 1) for k, v in r['phases_ms_max_over_ranks'].items()}, 'xor', r['checksum_xor_of_calls'], 'sum', r['checksum_sum_of_calls'])
"
done
