set -x
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
python tools/ab.py > $O/ab_main2.json 2> $O/ab_main2.err; cat $O/ab_main2.json; tail -2 $O/ab_main2.err
B200W_LIB=$PWD/build_variants/lib_j1mb20.so python tools/ab.py scat dtcwt > $O/ab_j1mb20.json 2> $O/ab_j1mb20.err; cat $O/ab_j1mb20.json; tail -2 $O/ab_j1mb20.err
python tools/policy_probe.py > $O/policy_main2.json 2>&1; cat $O/policy_main2.json
cap() { # name kernel-regex skip args...
  name=$1; rx=$2; skip=$3; shift 3
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx --launch-skip $skip -c 1 -f -o $O/r02b_$name python tools/prof_run.py "$@" > $O/r02b_ncu_$name.log 2>&1
  ncu -i $O/r02b_$name.ncu-rep --page details > $O/r02b_$name.details.txt 2>&1
  ncu -i $O/r02b_$name.ncu-rep --page raw --csv > $O/r02b_$name.raw.csv 2>&1
  tail -2 $O/r02b_ncu_$name.log
  rm -f $O/r02b_$name.ncu-rep
}
cap scat_l2 fwd_j1_stream 3 scat 256 2
cap pyramid_l1 dwt_pyramid 1 dwt 128 2
