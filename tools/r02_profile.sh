set -u
OUT=gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --aux off > $OUT/r02_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --aux off > $OUT/r02_ncu_launch.log 2>&1
bash tools/ncu_full.sh r02full 'conv3_fused|flash_d512|igemm_kernel|conv_in_kernel' 1 400
python tools/make_traffic_json.py $OUT/r02full_raw.csv $OUT/r02_dram_traffic.json
ls -la $OUT | tail -12
