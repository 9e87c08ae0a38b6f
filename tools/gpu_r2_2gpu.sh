set -x
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_parallel_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r02b_bench_n2.json 2> $O/r02b_bench_n2.err; tail -c 2500 $O/r02b_bench_n2.json; tail -3 $O/r02b_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c5 --steps 8 --warmup 3 > $O/r02b_bench_c5_n2.json 2> $O/r02b_bench_c5_n2.err; tail -c 1500 $O/r02b_bench_c5_n2.json; tail -3 $O/r02b_bench_c5_n2.err
