#!/bin/bash
# Round-2 first GPU checkpoint: parity tests, bf16-vs-oracle probe, bench line, reference arm.  Usage: tools/gpu_r2_first.sh TAG
TAG=${1:-r4a}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
python tools/bf16_parity_probe.py --json gpurun_out/bf16_probe_$TAG.json > gpurun_out/bf16_probe_$TAG.log 2>&1; echo "probe rc=$?"; tail -3 gpurun_out/bf16_probe_$TAG.log | cut -c1-400
python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/launch_table_$TAG.json > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err
python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print('value',d['value'],'e2e',d['e2e']['value'],'ms',d['ms_per_step'],'cpu',d['cpu_baseline'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_ref_$TAG.json
