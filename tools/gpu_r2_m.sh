set -x
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python tools/ab.py > $O/ab_main10.json 2> $O/ab_main10.err; cat $O/ab_main10.json; tail -2 $O/ab_main10.err
python tools/ab.py > $O/ab_main11.json 2> $O/ab_main11.err; cat $O/ab_main11.json
