set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 1500 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
timeout 200 python tools/bench_extra.py > gpurun_out/bench_extra.json 2>&1; tail -c 800 gpurun_out/bench_extra.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:afb2d_stream --launch-skip 3 -c 1 -f -o gpurun_out/afb_r1e python tools/prof_run.py dwt 128 2 > gpurun_out/ncu_afb.log 2>&1
ncu -i gpurun_out/afb_r1e.ncu-rep --page details > gpurun_out/afb_r1e.details.txt 2>&1
ncu -i gpurun_out/afb_r1e.ncu-rep --page raw --csv > gpurun_out/afb_r1e.raw.csv 2>&1
ls -la gpurun_out | tail -15
