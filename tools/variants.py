"""Build experimental variants of libb200wave.so (compile-time switches) into build_variants/ and, on a GPU box,
time the DWT / DTCWT forward with each:  python tools/variants.py build | run"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VARIANTS = {
    'default': [],
    'l2_256': ['-DB200W_CPASYNC_L2=256'],
    'l2_128': ['-DB200W_CPASYNC_L2=128'],
    'nostream': ['-DB200W_STREAM_STORES=0'],
}
OUT = os.path.join(ROOT, 'build_variants')

def build():
    from pytorch_wavelets_b200 import _build
    os.makedirs(OUT, exist_ok=True)
    for name, flags in VARIANTS.items():
        _build.build(out=os.path.join(OUT, 'lib_%s.so' % name), extra_flags=flags)
        print('built', name)

CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import _ffi
x = torch.randn(128, 32, 512, 512, device='cuda'); f = pw.DWTForward(J=3, wave='db4', mode='symmetric').cuda()
xt = torch.randn(64, 3, 1024, 1024, device='cuda'); d = pw.DTCWTForward(J=3).cuda()
res = {}
with torch.no_grad():
    for name, fn, inp in (('dwt', f, x), ('dtcwt', d, xt)):
        for _ in range(3): fn(inp)
        torch.cuda.synchronize()
        rec = _ffi.CallRecorder()
        with rec:
            for _ in range(10): fn(inp)
        s = rec.summary()
        res[name] = {k.split()[1]: round(v['avg_ms'], 4) for k, v in sorted(s.items())}
        res[name]['total'] = round(sum(v['avg_ms'] for v in s.values()), 4)
print(json.dumps(res))
'''

def run():
    for name in VARIANTS:
        so = os.path.join(OUT, 'lib_%s.so' % name)
        env = dict(os.environ, B200W_LIB=so)
        r = subprocess.run([sys.executable, '-c', CHILD % ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
        print(name, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])

if __name__ == '__main__':
    (build if sys.argv[1:] == ['build'] else run)()
