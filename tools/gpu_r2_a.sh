# round-2 run A: state check after the TU split + evidence the round-1 verdict asked for (sanitizer, ncu of the sub-0.5 kernels)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/bench_extra.py > $O/r02a_bench_extra.json 2>&1; tail -c 1200 $O/r02a_bench_extra.json
# compute-sanitizer over the small-shape tests (golden vectors + every-width sweeps are small)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "golden or periodization_full_depth or distinct_row_col or noncontiguous" > $O/r02a_memcheck.log 2>&1; echo memcheck rc=$?; tail -5 $O/r02a_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "dwt_golden or scat_golden or dtcwt_golden and (J3_64 or J2_40)" > $O/r02a_racecheck.log 2>&1; echo racecheck rc=$?; tail -5 $O/r02a_racecheck.log
prof() {  # name regex script-args...
  local name=$1 rx=$2; shift 2
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip 2 -c 1 -f -o $O/r02a_$name python tools/prof_run.py "$@" > $O/r02a_ncu_$name.log 2>&1
  ncu -i $O/r02a_$name.ncu-rep --page details > $O/r02a_$name.details.txt 2>&1
  ncu -i $O/r02a_$name.ncu-rep --page raw --csv > $O/r02a_$name.raw.csv 2>&1
}
prof afb16_c5 'afb2d_stream<16' c5 8 2
prof sfb8_c2 'sfb2d_stream<8' dwtinv 128 2
prof invj1_c3 inv_j1_stream dtcwtinv 64 2
prof invj2_c3 inv_j2plus_stream dtcwtinv 64 2
prof scat_c4 fwd_j1_stream scat 256 2
ls -la $O | tail -30
