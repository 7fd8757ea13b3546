set -x
timeout 600 python -m pytest tests/test_fused_features_gpu.py tests/test_tpconv_gpu.py -x -q 2>&1 | tail -15
DDB200_FUSED_DEBUG=1 timeout 300 python tools/bench_fused.py > gpurun_out/r02b_fused_dbg.json 2>&1
timeout 300 python tools/bench_fused.py > gpurun_out/r02b_fused.json 2>&1
timeout 300 python tools/bench_fused.py --layer 0 > gpurun_out/r02b_fused_l0.json 2>&1
DDB200_FUSED_CTA_PAIR=0 timeout 300 python tools/bench_fused.py > gpurun_out/r02b_fused_single.json 2>&1
cat gpurun_out/r02b_fused_dbg.json gpurun_out/r02b_fused.json gpurun_out/r02b_fused_l0.json gpurun_out/r02b_fused_single.json
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02b_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['achieved'],d['roofline']['kernel_ms_per_step'])"
