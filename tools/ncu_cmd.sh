#!/bin/bash
# `ncu --set full` of the kernels matching a regex in an arbitrary command (run once without ncu first), raw
# page exported to CSV on the GPU box:
#   tools/ncu_cmd.sh <tag> <kernel-regex> <count> <skip> -- <command ...>
set -u
TAG=$1; RX=$2; CNT=$3; SKIP=$4; shift 5
OUT=gpurun_out
mkdir -p $OUT
"$@" > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$RX" -s $SKIP -c $CNT -o /tmp/$TAG -f \
    "$@" > $OUT/ncu_$TAG.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
tail -1 $OUT/ncu_$TAG.log | cut -c1-200; wc -l $OUT/${TAG}_raw.csv
