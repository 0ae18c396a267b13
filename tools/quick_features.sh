#!/bin/bash
# Quick check on the GPU box: ORB / SIFT parity tests and the streaming / offline frames/s of BASELINE configs 3 and 4.
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_orb.py tests/test_gpu_pipeline.py tests/test_gpu_offline.py -q -m gpu -x -k "sift or feature_lock or featprep or orb" 2>&1 | tail -3
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-gpu 128 2>gpurun_out/quick_features.err | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print({k:(round(v['value'],1), v['matches'], v['inliers']) for k,v in d['other_modes_streaming'].items()})
        print({k:(round(v['value'],1), v.get('registrations_valid')) for k,v in d['other_modes_offline'].items()})
"
tail -3 gpurun_out/quick_features.err
