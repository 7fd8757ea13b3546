set -x
timeout 900 python -m pytest tests/test_sync_free_gpu.py -x -q 2>&1 | tail -30
DDB200_FUSED_DEBUG=1 timeout 300 python tools/bench_fused.py > gpurun_out/r02d_fused_dbg.json 2>&1
DDB200_FUSED_DEBUG=1 timeout 300 python tools/bench_fused.py --edges 1600000 > gpurun_out/r02d_fused_dbg_big.json 2>&1
cat gpurun_out/r02d_fused_dbg.json gpurun_out/r02d_fused_dbg_big.json
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err
tail -5 gpurun_out/r02d_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02d_bench.json'));print(d['value'],d['ms_per_step'],d['e2e'],d['roofline']['achieved'],d['roofline']['kernel_ms_per_step'],d['gpu_launches'])"
