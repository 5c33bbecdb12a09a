#!/bin/bash
# ncu --set full capture of microbench cases.  Usage: tools/gpu_ncu_mb.sh TAG mb-filter kernel-regex [count]
TAG=$1; FILT=$2; KRE=$3; CNT=${4:-2}
mkdir -p gpurun_out
python tools/microbench.py $FILT > gpurun_out/mb_$TAG.log 2>&1; echo "mb rc=$?"; cat gpurun_out/mb_$TAG.log
ncu --set full --import-source on --clock-control none -k regex:$KRE -c $CNT -f -o gpurun_out/prof_$TAG \
    python tools/microbench.py $FILT --once > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
