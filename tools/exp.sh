cd /root/repo
timeout 1500 python -m pytest tests/test_gpu_orb.py tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --frames-per-gpu 128 > gpurun_out/exp.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/exp.log").read().strip().splitlines()[-1])
print({k: round(v["value"],1) for k,v in d["other_modes_streaming"].items()})
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sift_launches.csv python tools/mode_probe.py sift 7 > gpurun_out/sift_probe.log 2>&1
