set -x
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_aa_model_gpu.py -m gpu -q -x --timeout 150 --timeout-method thread 2>&1 | tail -60 ) > gpurun_out/r02n_pytest_aa.txt 2>&1; cat gpurun_out/r02n_pytest_aa.txt | cut -c1-250
grep -q "failed\|error" gpurun_out/r02n_pytest_aa.txt && exit 1
( time timeout 600 python -m pytest tests/test_confidence_gpu.py tests/test_sync_free_gpu.py tests/test_sampler_gpu.py tests/test_model_gpu.py -m gpu -q -x --timeout 150 --timeout-method thread 2>&1 | tail -30 ) > gpurun_out/r02n_pytest_rest.txt 2>&1; cat gpurun_out/r02n_pytest_rest.txt | cut -c1-250
DDB200_CONFIG5_TRACE=1 timeout 300 python bench.py --workload config5 > gpurun_out/r02n_config5_trace.json 2> gpurun_out/r02n_config5_trace.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02n_config5_trace.json'))
print(d['value'], d['e2e']['seconds_per_run'])
tr=d['trace_rank0']
print([t[3] for t in tr])
print([round(t[1]*t[2]/1000) for t in tr])
PY
