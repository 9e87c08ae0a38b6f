set -x
O=gpurun_out
mkdir -p $O
for c in 0 1 2; do echo "=== case $c"; compute-sanitizer --tool racecheck build_variants/rc_probe $c 2>&1 | tail -12; done > $O/r02b_racecheck_mbarrier_probe.log 2>&1
cat $O/r02b_racecheck_mbarrier_probe.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_pyramid.py -m gpu -q -x -k "repeatability" 2>&1 | tail -3
