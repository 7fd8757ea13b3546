set -x
timeout 900 python -m pytest tests/test_aa_model_gpu.py tests/test_graph_gpu.py -q 2>&1 | tail -12
timeout 300 python tools/profile_sampling.py > gpurun_out/r02i_profile_small.json 2>gpurun_out/r02i_profile_small.err; tail -2 gpurun_out/r02i_profile_small.err
timeout 300 python tools/profile_sampling.py --n-res 1500 --n-atoms 40 --share 0 > gpurun_out/r02i_profile_big.json 2>gpurun_out/r02i_profile_big.err
cat gpurun_out/r02i_profile_small.json gpurun_out/r02i_profile_big.json
