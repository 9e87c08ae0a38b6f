set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pyramid.py -m gpu -q -x 2>&1 | tail -6
python tools/ab.py dwti dwt > $O/ab_main3.json 2> $O/ab_main3.err; cat $O/ab_main3.json; tail -2 $O/ab_main3.err
B200W_LIB=$PWD/build_variants/lib_base.so python tools/ab.py dwti > $O/ab_base3.json 2> $O/ab_base3.err; cat $O/ab_base3.json
python - <<'PY'
import torch, json, sys
sys.path.insert(0,'.')
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import _ffi
x = torch.randn(128, 32, 512, 512, device='cuda')
with torch.no_grad():
    f = pw.DWTForward(J=3, wave='db4', mode='symmetric').cuda(); g = pw.DWTInverse(wave='db4', mode='symmetric').cuda()
    c = f(x)
    for _ in range(3): g(c)
    with _ffi.CallRecorder() as rec:
        for _ in range(10): g(c)
    print(json.dumps({k: round(v['avg_ms'], 4) for k, v in rec.summary().items()}))
PY
