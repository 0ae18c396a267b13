cd /root/repo
timeout 1200 python -m pytest tests/test_gpu_orb.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python - <<PY
import json, subprocess, sys
sys.argv=["bench.py"]
PY
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-gpu 128 > gpurun_out/exp.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/exp.log").read().strip().splitlines()[-1])
print(json.dumps(d["other_modes_streaming"], indent=1))
PY
