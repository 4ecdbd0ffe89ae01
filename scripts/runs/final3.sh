set -x
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_final3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_final3.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
