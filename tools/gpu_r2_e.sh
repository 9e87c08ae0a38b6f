set -x
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -2
run() { # name lib args...
  n=$1; l=$2; shift 2
  if [ "$l" = default ]; then python tools/ab.py "$@" > $O/ab_$n.json 2> $O/ab_$n.err; else B200W_LIB=$PWD/$l python tools/ab.py "$@" > $O/ab_$n.json 2> $O/ab_$n.err; fi
  cat $O/ab_$n.json; tail -2 $O/ab_$n.err
}
run default build_variants/lib_base.so dwt scat dtcwt
run s4 build_variants/lib_s4.so dwt
run s4b build_variants/lib_s4b.so dwt
run j1new build_variants/lib_j1new.so scat dtcwt
run j1mb20 build_variants/lib_j1mb20.so scat dtcwt
run default2 build_variants/lib_base.so dwt scat dtcwt
B200W_LIB=$PWD/build_variants/lib_s4.so timeout 300 python tools/pyr_check.py check 2>&1 | tail -4
B200W_LIB=$PWD/build_variants/lib_j1new.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scat or dtcwt" 2>&1 | tail -4
