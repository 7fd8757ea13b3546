set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_conv_kernel -s 2 -c 1 -o gpurun_out/r02m_fused python tools/bench_fused.py > gpurun_out/r02m_ncu_fused.log 2>&1
ncu -i gpurun_out/r02m_fused.ncu-rep --page raw --csv > gpurun_out/r02m_fused_raw.csv 2>/dev/null
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02m_launches.csv python bench.py --steps 2 --warmup 3 --short-warmup --no-e2e --quick > gpurun_out/r02m_ncu_bench.log 2>&1
tail -3 gpurun_out/r02m_ncu_bench.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r02m_clocks.csv &
SMI=$!
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
kill $SMI
tail -3 gpurun_out/r02m_bench.err; cut -c1-200 gpurun_out/r02m_bench.json
timeout 600 python bench.py --workload config5 > gpurun_out/r02m_config5_n1.json 2> gpurun_out/r02m_config5_n1.err; cut -c1-300 gpurun_out/r02m_config5_n1.json
timeout 600 python tools/bench_fused.py --scan > gpurun_out/r02m_config4_fused.jsonl 2>&1; tail -2 gpurun_out/r02m_config4_fused.jsonl
ls -la gpurun_out | tail -12
