#!/bin/bash
# A/B of the LK kernel launch shapes on the GPU box: prints the stage times per combination.
cd "$(dirname "$0")/.."
ARGS="--steps 6 --warmup 3 --frames-per-gpu 256 --batch 256 --no-cpu-baseline --no-e2e --no-mode-probes"
run() {
  echo "== $*"
  env "$@" python bench.py $ARGS 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); s=d['stages']
        print('value %.0f ms/step %.3f' % (d['value'], d['ms_per_step']), ' '.join('%s %.3f' % (k, v['ms_per_step']) for k, v in s.items()))
"
}
for r in 80 72 64; do run VSTAB_LK_WARPS=1 VSTAB_LK_REGS=$r; done
