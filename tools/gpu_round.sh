#!/bin/bash
# One GPU-box round: parity tests, bench line (+ per-launch table), per-op microbench.  Usage: tools/gpu_round.sh TAG [mb-filter]
TAG=${1:-x}; FILT=${2:-}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/launch_table_$TAG.json > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'])"
python tools/microbench.py $FILT > gpurun_out/mb_$TAG.log 2>&1; echo "mb rc=$?"; cat gpurun_out/mb_$TAG.log
