# round-2b final evidence run (one B200): tests, smoke, bench lines of all three arms, ncu launch list + full captures,
# compute-sanitizer over the small-shape tests.  Outputs under gpurun_out/r02b_*; the judged copies go to profiles/.
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02b_bench_n1.json 2> $O/r02b_bench_n1.err; tail -c 600 $O/r02b_bench_n1.json; tail -2 $O/r02b_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02b_bench_reference_arm.json 2> $O/r02b_bench_reference_arm.err; tail -c 500 $O/r02b_bench_reference_arm.json
timeout 600 python bench.py --impl aten --steps 5 --warmup 3 > $O/r02b_bench_aten_arm.json 2> $O/r02b_bench_aten_arm.err; tail -c 400 $O/r02b_bench_aten_arm.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parts > $O/r02b_bench_under_ncu.log 2>&1
cap() { # name kernel-regex skip args...
  name=$1; rx=$2; skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 -f -o $O/r02b_$name python tools/prof_run.py "$@" > $O/r02b_ncu_$name.log 2>&1
  ncu -i $O/r02b_$name.ncu-rep --page details > $O/r02b_ncu_full_$name.details.txt 2>&1
  ncu -i $O/r02b_$name.ncu-rep --page raw --csv > $O/r02b_ncu_full_$name.raw.csv 2>&1
  tail -1 $O/r02b_ncu_$name.log
  rm -f $O/r02b_$name.ncu-rep
}
cap pyramid_l1 dwt_pyramid 1 dwt 128 2
cap sfb4_c2 sfb2d_stream4 5 dwtinv 128 2
cap scat_l2 fwd_j1_stream 3 scat 256 2
cap afb16_c5 afb2d_stream 4 c5 8 2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_prims.py tests/test_gpu_f64.py -x -q -m gpu -k "golden or periodization_full_depth or distinct_row_col or noncontiguous or wide or prims or f64 or other_families" > $O/r02b_sanitizer_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 $O/r02b_sanitizer_memcheck.log
# racecheck over the kernels that synchronise with __syncwarp / __syncthreads / cp.async groups (what the tool models).  The
# pyramid kernel's mbarrier + TMA complete_tx ordering is not modelled (every producer -> consumer pair is flagged: excerpt
# in profiles/r02b_sanitizer_racecheck_pyramid_excerpt.log, discussion in profiles/r02_notes.md), so its tests are left out.
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_prims.py -x -q -m gpu -k "dwt_golden or scat_golden or (wide_synthesis and db2) or (dtcwt_golden and (J3_64 or J2_40)) or (prims and qshift_a)" > $O/r02b_sanitizer_racecheck.log 2>&1; echo racecheck rc=$?; tail -4 $O/r02b_sanitizer_racecheck.log
ls -la $O | grep r02b_ | head -40
