#!/bin/bash
# Round evidence on one B200: parity tests, the bench line (with cpu_baseline), the per-launch table of one step, then —
# only after the plain run exited 0 — the ncu launch list (time + DRAM bytes per launch) of one profiled step, joined
# with the plan tags, and the per-instance DRAM traffic table bench.py reads (profiles/traffic.json).
# Usage: tools/gpu_evidence.sh TAG [skip-tests]
TAG=${1:-x}
mkdir -p gpurun_out
if [ -z "$2" ]; then
  python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_$TAG.log
fi
python bench.py --steps 20 --warmup 3 --profile-out gpurun_out/launch_table_$TAG.json > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
rc=$?; echo "bench rc=$rc"; tail -3 gpurun_out/bench_$TAG.err
python -c "import json;d=json.load(open('gpurun_out/bench_$TAG.json'));print('value',d['value'],'e2e',d['e2e']['value'],'e2e_fp32',(d.get('e2e_fp32_feed') or {}).get('value'),'roof',{k:(v['instance'],round(v['frac'],3)) for k,v in d['roofline']['classes'].items()},'whole',d['roofline']['whole_step']['frac_of_bf16_sustained'],'cpu',d['cpu_baseline']['value'])"
if [ $rc -eq 0 ]; then
  python bench.py --profile-step > gpurun_out/plain_profile_step_$TAG.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
      --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --profile-step > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
  python tools/join_ncu_tags.py gpurun_out/launches_$TAG.csv gpurun_out/launch_table_$TAG.json > gpurun_out/launches_by_tag_$TAG.txt; head -14 gpurun_out/launches_by_tag_$TAG.txt
  python tools/ncu_traffic.py gpurun_out/launches_$TAG.csv gpurun_out/launch_table_$TAG.json gpurun_out/traffic_$TAG.json | head -8
fi
