set -x
mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_sync_free_gpu.py tests/test_sampler_gpu.py tests/test_aa_model_gpu.py tests/test_inputs_gpu.py tests/test_confidence_gpu.py tests/test_pyg_like_gpu.py -m gpu -q -x --timeout 150 --timeout-method thread 2>&1 | tail -40 ) > gpurun_out/r02o_pytest.txt 2>&1; cat gpurun_out/r02o_pytest.txt | cut -c1-250
grep -q "failed\|error" gpurun_out/r02o_pytest.txt && exit 1
timeout 200 python tools/profile_config5.py --n 24 --order index > gpurun_out/r02o_cfg5_anat_index.jsonl 2>&1; tail -26 gpurun_out/r02o_cfg5_anat_index.jsonl | cut -c1-200
timeout 200 python tools/profile_config5.py --n 24 --order index --sync 0 > gpurun_out/r02o_cfg5_index_nosync.jsonl 2>&1; tail -1 gpurun_out/r02o_cfg5_index_nosync.jsonl
timeout 300 python bench.py --workload config5 > gpurun_out/r02o_config5_n1.json 2> gpurun_out/r02o_config5_n1.err; cut -c1-200 gpurun_out/r02o_config5_n1.json
( time timeout 900 python bench.py ) > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err; tail -3 gpurun_out/r02o_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02o_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['runs_s'], d['config2_batch32'], d['cfg_l1_sh_lmax1'])
PY
