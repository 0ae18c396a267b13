cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_orb.py tests/test_gpu_pipeline.py tests/test_gpu_offline.py -q -m gpu -x 2>&1 | tail -3
for g in 1 0; do
VSTAB_LOOKAHEAD=$g timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-gpu 128 > gpurun_out/exp.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/exp.log").read().strip().splitlines()[-1])
print("lookahead $g", {k: (round(v["value"],1), v["matches"], v["inliers"]) for k,v in d["other_modes_streaming"].items()})
PY
done
