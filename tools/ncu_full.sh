#!/bin/bash
# One `ncu --set full` capture of the hot kernels of a single encode+tag pass; raw + source pages
# exported to CSV on the GPU box (the .ncu-rep itself is kept only when it is small enough to travel).
#   tools/ncu_full.sh <tag> <kernel-regex> [batch] [max kernels]
set -u
TAG=${1:-full}; RX=${2:-conv3_fused|flash_d512}; B=${3:-1}; CNT=${4:-30}
OUT=gpurun_out
mkdir -p $OUT
python tools/one_pass.py $B 1024 1 > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RX" -c $CNT -o /tmp/$TAG -f \
    python tools/one_pass.py $B 1024 1 > $OUT/ncu_$TAG.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > $OUT/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/$TAG.ncu-rep --page details --csv > $OUT/${TAG}_details.csv 2>/dev/null
ncu -i /tmp/$TAG.ncu-rep --page source --csv -k regex:flash_d512 > $OUT/${TAG}_src_flash.csv 2>/dev/null
ncu -i /tmp/$TAG.ncu-rep --page source --csv -k regex:"conv3_fused_kernel<128" -c 1 > $OUT/${TAG}_src_conv3t.csv 2>/dev/null
SZ=$(stat -c %s /tmp/$TAG.ncu-rep)
echo "rep size $SZ"
if [ "$SZ" -lt 40000000 ]; then cp /tmp/$TAG.ncu-rep $OUT/; fi
ls -la $OUT | tail -8
