set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_fused.py > gpurun_out/r02l_fused.json 2>&1; cat gpurun_out/r02l_fused.json
DDB200_FUSED_DEBUG=1 timeout 300 python tools/bench_fused.py > gpurun_out/r02l_fused_dbg.json 2>&1; cat gpurun_out/r02l_fused_dbg.json
( time timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > gpurun_out/r02l_pytest_gpu.txt 2>&1; cat gpurun_out/r02l_pytest_gpu.txt
timeout 300 python tools/bench_fused.py --edges 1600000 --nodes 70000 > gpurun_out/r02l_fused_1m6.json 2>&1; cat gpurun_out/r02l_fused_1m6.json
( time timeout 900 python bench.py ) > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; tail -4 gpurun_out/r02l_bench.err; cut -c1-300 gpurun_out/r02l_bench.json
