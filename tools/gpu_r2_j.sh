set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "wide or dtcwt" 2>&1 | tail -6
python tools/ab.py dtcwti c5 > $O/ab_main5.json 2> $O/ab_main5.err; cat $O/ab_main5.json; tail -2 $O/ab_main5.err
for v in invw_mb12 invw_mb14 base; do B200W_LIB=$PWD/build_variants/lib_$v.so python tools/ab.py dtcwti > $O/ab_$v.json 2> $O/ab_$v.err; cat $O/ab_$v.json; tail -1 $O/ab_$v.err; done
B200W_LIB=$PWD/build_variants/lib_afb16_mb14.so python tools/ab.py c5 > $O/ab_afb16_mb14.json 2> $O/ab_afb16_mb14.err; cat $O/ab_afb16_mb14.json
