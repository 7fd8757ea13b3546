set -x
timeout 600 python -m pytest tests/test_fused_features_gpu.py tests/test_tpconv_gpu.py -x -q 2>&1 | tail -5
DDB200_FUSED_DEBUG=1 timeout 300 python tools/bench_fused.py > gpurun_out/r02c_fused_dbg.json 2>&1
DDB200_FUSED_DEBUG=1 DDB200_FUSED_NOLOAD=1 timeout 300 python tools/bench_fused.py > gpurun_out/r02c_fused_dbg_noload.json 2>&1
timeout 300 python tools/bench_fused.py > gpurun_out/r02c_fused.json 2>&1
timeout 300 python tools/bench_fused.py --layer 0 > gpurun_out/r02c_fused_l0.json 2>&1
DDB200_FUSED_CTA_PAIR=0 timeout 300 python tools/bench_fused.py > gpurun_out/r02c_fused_single.json 2>&1
cat gpurun_out/r02c_fused_dbg.json gpurun_out/r02c_fused_dbg_noload.json gpurun_out/r02c_fused.json gpurun_out/r02c_fused_l0.json gpurun_out/r02c_fused_single.json
