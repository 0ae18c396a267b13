cd /root/repo
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -q -m gpu -x 2>&1 | tail -8
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-mode-probes > gpurun_out/exp.log 2>&1
python - <<PY
import json,sys
d=json.loads(open("gpurun_out/exp.log").read().strip().splitlines()[-1])
s=d["stages"]; n=d["config"]["frames_per_gpu"]
print("value", round(d["value"]), {k: round(v["ms_per_step"],3) for k,v in s.items()})
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'warp_|eig_kernel|greedy|ingest|lkprep' -c 5 -o gpurun_out/prof_r1b -f python bench.py --steps 1 --warmup 1 --frames-per-gpu 128 --no-cpu-baseline --no-e2e --no-mode-probes > gpurun_out/ncu_r1b.log 2>&1
ls -la gpurun_out/*.ncu-rep
