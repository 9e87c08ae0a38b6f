# round-2 run B: tests, smoke, the three bench arms
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; tail -c 6000 $O/r02_bench_n1.json; tail -5 $O/r02_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.json 2> $O/r02_bench_ref.err; tail -c 1200 $O/r02_bench_ref.json
timeout 600 python bench.py --impl aten --steps 5 --warmup 2 > $O/r02_bench_aten.json 2> $O/r02_bench_aten.err; tail -c 900 $O/r02_bench_aten.json; tail -3 $O/r02_bench_aten.err
