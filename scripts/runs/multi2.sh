set -x
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 --cpu-frames 300 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err
timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -x -q > gpurun_out/t_shard.log 2>&1; echo "rc=$?" >> gpurun_out/t_shard.log
