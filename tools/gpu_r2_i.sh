set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide or inverse or golden or oracle" 2>&1 | tail -4
python tools/ab.py dwti c5 > $O/ab_main4.json 2> $O/ab_main4.err; cat $O/ab_main4.json; tail -2 $O/ab_main4.err
for v in afb16_mb16 afb16_mb12; do B200W_LIB=$PWD/build_variants/lib_$v.so python tools/ab.py c5 > $O/ab_$v.json 2> $O/ab_$v.err; cat $O/ab_$v.json; tail -1 $O/ab_$v.err; done
