set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 2 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err
