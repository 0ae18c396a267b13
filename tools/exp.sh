cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k warp 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-mode-probes --no-e2e > gpurun_out/exp.log 2>&1
python - <<PY
import json,sys
d=json.loads(open("gpurun_out/exp.log").read().strip().splitlines()[-1])
s=d["stages"]
print("value", round(d["value"]), {k: round(v["ms_per_step"],3) for k,v in s.items()})
PY
done
