set -x
O=gpurun_out
mkdir -p $O
echo default; timeout 200 python tools/policy_probe.py 2>&1 | tail -1
echo fuseall; B200W_LIB=build_variants/fuseall.so timeout 200 python tools/policy_probe.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-parts > $O/r02_bench_under_ncu.log 2>&1
prof() {  # name regex skip script-args...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 -f -o $O/r02_$name python tools/prof_run.py "$@" > $O/r02_ncu_$name.log 2>&1
  ncu -i $O/r02_$name.ncu-rep --page details > $O/r02_$name.details.txt 2>&1
  ncu -i $O/r02_$name.ncu-rep --page raw --csv > $O/r02_$name.raw.csv 2>&1
  tail -2 $O/r02_ncu_$name.log
}
prof pyramid_l1 dwt_pyramid 1 dwt 128 2
prof afb16_c5 afb2d_stream 4 c5 8 2
prof sfb8_c2 sfb2d_stream 5 dwtinv 128 2
prof invj1_c3 inv_j1_stream 1 dtcwtinv 64 2
prof afb8_l2 afb2d_stream 2 dwt 128 2
ls -la $O | grep r02_ | head -40
