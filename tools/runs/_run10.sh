set -x
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02h_bench_n$N.json 2> gpurun_out/r02h_bench_n$N.err
tail -3 gpurun_out/r02h_bench_n$N.err; cut -c1-400 gpurun_out/r02h_bench_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --workload config5 > gpurun_out/r02h_config5_n$N.json 2> gpurun_out/r02h_config5_n$N.err
tail -3 gpurun_out/r02h_config5_n$N.err; cut -c1-300 gpurun_out/r02h_config5_n$N.json
