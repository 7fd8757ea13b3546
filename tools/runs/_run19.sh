set -x
mkdir -p gpurun_out
timeout 200 python tools/profile_config5.py --n 24 --order index > gpurun_out/r02n_cfg5_anat_index.jsonl 2>&1; tail -26 gpurun_out/r02n_cfg5_anat_index.jsonl
timeout 200 python tools/profile_config5.py --n 24 --order desc --sync 0 > gpurun_out/r02n_cfg5_desc_nosync.jsonl 2>&1; tail -1 gpurun_out/r02n_cfg5_desc_nosync.jsonl
timeout 200 python tools/profile_config5.py --n 24 --order index --sync 0 > gpurun_out/r02n_cfg5_index_nosync.jsonl 2>&1; tail -1 gpurun_out/r02n_cfg5_index_nosync.jsonl
