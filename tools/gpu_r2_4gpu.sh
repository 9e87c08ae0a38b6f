set -x
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m 2>&1 | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r02b_bench_n4.json 2> $O/r02b_bench_n4.err; tail -c 2000 $O/r02b_bench_n4.json; tail -3 $O/r02b_bench_n4.err
