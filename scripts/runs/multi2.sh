set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 3 --multi gather > gpurun_out/bench_n2_gather.json 2> gpurun_out/bench_n2_gather.err
python bench.py --frames 2368 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n1_check.json 2> gpurun_out/bench_n1_check.err
