#!/bin/bash
# Multi-GPU evidence on one box (an N-GPU call is charged N x its wall time: every step below has its OWN short timeout —
# in r4 a train run whose ranks hung at exit held 8 GPUs for its whole 600 s limit and spent the round's budget): inference bench (one global EDM threshold: exit counts differ per rank) and the config-3
# train step at N ranks.  Usage: tools/gpu_multi.sh N TAG
N=${1:-2}; TAG=${2:-x}
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_${N}gpu_$TAG.json'));print('value',d['value'],'e2e',d['e2e']['value'],'fp32feed',(d.get('e2e_fp32_feed') or {}).get('value'),'ranks',d['ranks'])"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 3 --per-rank-threshold --no-fp32-feed > gpurun_out/bench_${N}gpu_perrank_$TAG.json 2> gpurun_out/bench_${N}gpu_perrank_$TAG.err; echo "bench(per-rank thr) rc=$?"
python -c "import json;d=json.load(open('gpurun_out/bench_${N}gpu_perrank_$TAG.json'));print('value',d['value'],'e2e',d['e2e']['value'])"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 tools/train_bench.py --steps 3 --warmup 3 --check > gpurun_out/train_${N}gpu_$TAG.json 2> gpurun_out/train_${N}gpu_$TAG.err; echo "train rc=$?"
tail -1 gpurun_out/train_${N}gpu_$TAG.json | cut -c1-1400
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 tools/train_bench.py --steps 3 --warmup 3 --exchange nccl > gpurun_out/train_${N}gpu_nccl_$TAG.json 2> gpurun_out/train_${N}gpu_nccl_$TAG.err; echo "train(nccl exchange) rc=$?"
tail -1 gpurun_out/train_${N}gpu_nccl_$TAG.json | cut -c1-300
