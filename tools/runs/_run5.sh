set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -3 gpurun_out/r02e_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02e_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['runs_s'],d.get('parity'),d.get('cpu_baseline'))"
timeout 900 python bench.py --workload config5 --complexes 16 > gpurun_out/r02e_config5_16.json 2> gpurun_out/r02e_config5.err
tail -3 gpurun_out/r02e_config5.err; cat gpurun_out/r02e_config5_16.json | cut -c1-1500
