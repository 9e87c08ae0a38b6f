set -x
O=gpurun_out
mkdir -p $O
python tools/ab.py dwt dwti dtcwt dtcwti > $O/ab_main6.json 2> $O/ab_main6.err; cat $O/ab_main6.json; tail -2 $O/ab_main6.err
B200W_LIB=$PWD/build_variants/lib_invj2_mb16.so python tools/ab.py dtcwti > $O/ab_k1.json 2>$O/ab_k1.err; cat $O/ab_k1.json
B200W_LIB=$PWD/build_variants/lib_fwdj2_mb16.so python tools/ab.py dtcwt > $O/ab_k2.json 2>$O/ab_k2.err; cat $O/ab_k2.json
B200W_LIB=$PWD/build_variants/lib_sfb4_ns2.so python tools/ab.py dwti > $O/ab_k3.json 2>$O/ab_k3.err; cat $O/ab_k3.json
B200W_LIB=$PWD/build_variants/lib_afb8_mb20.so python tools/ab.py dwt > $O/ab_k4.json 2>$O/ab_k4.err; cat $O/ab_k4.json
python tools/ab.py dwt dwti dtcwt dtcwti > $O/ab_main7.json 2> $O/ab_main7.err; cat $O/ab_main7.json
