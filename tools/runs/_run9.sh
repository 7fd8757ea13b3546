set -x
timeout 1200 python -m pytest tests/test_graph_gpu.py tests/test_sync_free_gpu.py tests/test_radial_gemm_gpu.py tests/test_confidence_gpu.py -q 2>&1 | tail -15
timeout 300 python tools/profile_sampling.py > gpurun_out/r02g_profile_small.json 2>gpurun_out/r02g_profile_small.err; tail -3 gpurun_out/r02g_profile_small.err
timeout 300 python tools/profile_sampling.py --n-res 1500 --n-atoms 40 --share 0 > gpurun_out/r02g_profile_big.json 2>gpurun_out/r02g_profile_big.err
cat gpurun_out/r02g_profile_small.json gpurun_out/r02g_profile_big.json
timeout 600 python tools/bench_tpconv.py --scan > gpurun_out/r02g_config4_tpconv.jsonl 2>&1; tail -2 gpurun_out/r02g_config4_tpconv.jsonl
timeout 900 python bench.py --workload config5 > gpurun_out/r02g_config5_n1.json 2> gpurun_out/r02g_config5_n1.err
tail -3 gpurun_out/r02g_config5_n1.err; cut -c1-300 gpurun_out/r02g_config5_n1.json
