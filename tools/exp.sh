cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-mode-probes > gpurun_out/exp.log 2>&1
python - <<PY
import json,sys
d=json.loads(open("gpurun_out/exp.log").read().strip().splitlines()[-1])
s=d["stages"]; n=d["config"]["frames_per_gpu"]
print("value", round(d["value"]), {k: round(v["ms_per_step"],3) for k,v in s.items()})
print("e2e", d["e2e"]["value"], "streaming", d["e2e"]["streaming"]["value"])
PY
