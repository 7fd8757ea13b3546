set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02f_pytest_gpu.txt; cat gpurun_out/r02f_pytest_gpu.txt
timeout 300 python tools/profile_sampling.py > gpurun_out/r02f_profile_small.json 2>gpurun_out/r02f_profile_small.err
timeout 300 python tools/profile_sampling.py --n-res 1500 --n-atoms 40 --share 0 > gpurun_out/r02f_profile_big.json 2>gpurun_out/r02f_profile_big.err
cat gpurun_out/r02f_profile_small.json gpurun_out/r02f_profile_big.json
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r02f_clocks.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err
kill $SMI
tail -3 gpurun_out/r02f_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02f_reference.json 2> gpurun_out/r02f_reference.err
cut -c1-900 gpurun_out/r02f_reference.json
timeout 600 python tools/bench_fused.py --scan > gpurun_out/r02f_config4_fused.jsonl 2>&1
timeout 600 python tools/bench_tpconv.py --scan > gpurun_out/r02f_config4_tpconv.jsonl 2>&1
tail -3 gpurun_out/r02f_config4_fused.jsonl gpurun_out/r02f_config4_tpconv.jsonl
timeout 900 python bench.py --workload config5 > gpurun_out/r02f_config5_n1.json 2> gpurun_out/r02f_config5_n1.err
cut -c1-400 gpurun_out/r02f_config5_n1.json
# ---- ncu: launch list of a short bench run, full captures of the two conv kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --short-warmup --no-e2e --quick > gpurun_out/r02f_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_conv_kernel -s 2 -c 1 -o gpurun_out/r02f_fused python tools/bench_fused.py > gpurun_out/r02f_ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tpconv_accumulate_kernel -s 2 -c 1 -o gpurun_out/r02f_tpconv python tools/bench_tpconv.py > gpurun_out/r02f_ncu_tpconv.log 2>&1
ncu -i gpurun_out/r02f_fused.ncu-rep --page raw --csv > gpurun_out/r02f_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/r02f_tpconv.ncu-rep --page raw --csv > gpurun_out/r02f_tpconv_raw.csv 2>/dev/null
ls -la gpurun_out | tail -30
