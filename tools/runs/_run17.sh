set -x
mkdir -p gpurun_out
N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29501 bench.py --gpus $N --workload config5 > gpurun_out/r02m_config5_n$N.json 2> gpurun_out/r02m_config5_n$N.err
tail -3 gpurun_out/r02m_config5_n$N.err; cut -c1-400 gpurun_out/r02m_config5_n$N.json
