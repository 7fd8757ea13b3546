set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_fused_features_gpu.py tests/test_tpconv_gpu.py -m gpu -q -x --timeout 120 --timeout-method thread 2>&1 | tail -5 > gpurun_out/r02m_pytest_first.txt; cat gpurun_out/r02m_pytest_first.txt
grep -q "passed" gpurun_out/r02m_pytest_first.txt && ! grep -q "failed\|Timeout\|error" gpurun_out/r02m_pytest_first.txt || exit 1
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 180 --timeout-method thread 2>&1 | tail -15 ) > gpurun_out/r02m_pytest_gpu.txt 2>&1; cat gpurun_out/r02m_pytest_gpu.txt
timeout 300 python tools/bench_inputs.py > gpurun_out/r02m_inputs.jsonl 2>&1; cat gpurun_out/r02m_inputs.jsonl
timeout 300 python tools/bench_fused.py > gpurun_out/r02m_fused.json 2>&1; cat gpurun_out/r02m_fused.json
