#!/bin/bash
# bench.py under a list of env settings (A/B runs of tuning knobs).  Usage: tools/gpu_sweep_env.sh "VAR=a" "VAR=b" ...
mkdir -p gpurun_out
for V in "$@"; do
  echo -n "== $V : "
  env $V timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/sweep.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1))" || (grep -v "^frame" gpurun_out/sweep.err | tail -2)
done
