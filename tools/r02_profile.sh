set -u
OUT=gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --aux off > $OUT/r02_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --aux off > $OUT/r02_ncu_launch.log 2>&1
bash tools/ncu_full.sh r02full 'conv3_fused|flash_d512|igemm_kernel|conv_in_kernel' 1 400
python tools/make_traffic_json.py $OUT/r02full_raw.csv $OUT/r02_dram_traffic.json
ls -la $OUT | tail -12
# the VAE fine-tuning pass: timed line + ncu launch list
python tools/encoder_train_bench.py --batch 2 --res 1024 --steps 5 > $OUT/r02_encoder_train_b2_1024.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/r02_launches_encoder_train.csv \
    python tools/encoder_train_bench.py --batch 2 --res 1024 --steps 1 --warmup 0 > /dev/null 2>&1
