cd /root/repo
VSTAB_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-mode-probes --frames-per-gpu 128 > gpurun_out/exp.log 2> gpurun_out/exp.err
grep "vstab trace" gpurun_out/exp.err
for i in 1 2; do
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-mode-probes --frames-per-gpu 128 > gpurun_out/exp.log 2>&1
python - <<PY
import json,sys
d=json.loads(open("gpurun_out/exp.log").read().strip().splitlines()[-1])
print("e2e", d["e2e"]["value"], "streaming", d["e2e"]["streaming"]["value"])
PY
done
