set -x
timeout 1200 python -m pytest tests/test_sync_free_gpu.py tests/test_sampler_gpu.py tests/test_model_gpu.py tests/test_full_size_gpu.py -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
tail -3 gpurun_out/r02k_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r02k_bench.json'));print(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['runs_s'],d['config2_batch32'],d['cfg_l1_sh_lmax1'],d['roofline']['frac'],d['roofline']['tensor']['issued_TFLOPs'])"
