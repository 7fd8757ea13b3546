DDB200_CONFIG5_TRACE=1 timeout 900 python bench.py --workload config5 > gpurun_out/r02j_config5_trace.json 2> gpurun_out/r02j_config5_trace.err
tail -2 gpurun_out/r02j_config5_trace.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02j_config5_trace.json'))
print(d['value'])
tr=d['trace_rank0']
print([t[3] for t in tr])
print([t[1]*t[2] for t in tr])
PY
