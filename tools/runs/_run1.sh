set -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
DDB200_FUSED_DEBUG=1 python tools/bench_fused.py > gpurun_out/r02a_fused_dbg.json 2>&1
python tools/bench_fused.py > gpurun_out/r02a_fused.json 2>&1
DDB200_FUSED_DEBUG=1 python tools/bench_fused.py --layer 0 > gpurun_out/r02a_fused_dbg_l0.json 2>&1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
cat gpurun_out/r02a_fused_dbg.json gpurun_out/r02a_fused.json gpurun_out/r02a_fused_dbg_l0.json
cat gpurun_out/r02a_bench.json | head -c 3000
