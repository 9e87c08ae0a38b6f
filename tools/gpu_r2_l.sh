set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide or inverse or golden or vs_oracle" 2>&1 | tail -3
python tools/ab.py dwti > $O/ab_main8.json 2> $O/ab_main8.err; cat $O/ab_main8.json; tail -2 $O/ab_main8.err
B200W_LIB=$PWD/build_variants/lib_sfb4_mb20.so python tools/ab.py dwti > $O/ab_l1.json 2>$O/ab_l1.err; cat $O/ab_l1.json
python tools/ab.py dwti > $O/ab_main9.json 2> $O/ab_main9.err; cat $O/ab_main9.json
