set -x
timeout 900 python -m pytest tests/test_confidence_gpu.py -x -q 2>&1 | tail -8
timeout 300 python tools/profile_sampling.py 2>&1 | tail -2
timeout 300 python tools/profile_sampling.py --n-res 1500 --n-atoms 40 --share 0 2>&1 | tail -2
