cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_orb.py tests/test_gpu_offline.py tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/exp.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/exp.log").read().strip().splitlines()[-1])
s=d["stages"]
print("value", round(d["value"]), {k: round(v["ms_per_step"],3) for k,v in s.items()})
print({k: round(v["value"],1) for k,v in d["other_modes_streaming"].items()})
PY
