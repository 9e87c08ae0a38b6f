# final bench lines of the shipped library + the ScatLayer capture of the final epilogue (after tools/final_run.sh)
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err; tail -c 300 $O/r02b_bench_n1.json; tail -2 $O/r02b_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02b_bench_reference_arm.json 2> $O/r02b_bench_reference_arm.err; tail -c 300 $O/r02b_bench_reference_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parts > $O/r02b_bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fwd_j1_stream --launch-skip 3 -c 1 -f -o $O/r02b_scat_l2 python tools/prof_run.py scat 256 2 > $O/r02b_ncu_scat_l2.log 2>&1
ncu -i $O/r02b_scat_l2.ncu-rep --page details > $O/r02b_ncu_full_scat_l2.details.txt 2>&1
ncu -i $O/r02b_scat_l2.ncu-rep --page raw --csv > $O/r02b_ncu_full_scat_l2.raw.csv 2>&1
rm -f $O/r02b_scat_l2.ncu-rep
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scat_golden or scat_magbias or scat_vs_oracle" > $O/r02b_sanitizer_memcheck_scat.log 2>&1; echo memcheck rc=$?; tail -3 $O/r02b_sanitizer_memcheck_scat.log
