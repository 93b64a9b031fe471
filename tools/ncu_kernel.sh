#!/bin/bash
# `ncu --set full` of a few launches of ONE kernel family in a single encode+tag pass; the report is
# small enough to travel back (source + raw pages are read locally with ncu -i).
#   tools/ncu_kernel.sh <tag> <kernel-regex> [batch] [count] [skip]
set -u
TAG=${1:-k}; RX=${2:-conv3_fused}; B=${3:-1}; CNT=${4:-2}; SKIP=${5:-0}
OUT=gpurun_out
mkdir -p $OUT
python tools/one_pass.py $B 1024 1 > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -o $OUT/$TAG -f \
    python tools/one_pass.py $B 1024 1 > $OUT/ncu_$TAG.log 2>&1
tail -2 $OUT/ncu_$TAG.log; ls -la $OUT/$TAG.ncu-rep
