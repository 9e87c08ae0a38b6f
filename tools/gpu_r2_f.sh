set -x
O=gpurun_out
mkdir -p $O
nvidia-smi -L | head -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/ab.py > $O/ab_main.json 2> $O/ab_main.err; cat $O/ab_main.json; tail -2 $O/ab_main.err
B200W_LIB=$PWD/build_variants/lib_base.so python tools/ab.py > $O/ab_base.json 2> $O/ab_base.err; cat $O/ab_base.json; tail -2 $O/ab_base.err
python tools/policy_probe.py > $O/policy_main.json 2>&1; cat $O/policy_main.json
B200W_LIB=$PWD/build_variants/lib_base.so python tools/policy_probe.py > $O/policy_base.json 2>&1; cat $O/policy_base.json
