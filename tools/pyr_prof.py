"""Wait-time breakdown of the fused pyramid kernel: runs a variant library built with -DB200W_PYR_PROF
(B200W_LIB=<path> python tools/pyr_prof.py).  Counters: cycles summed over warps (lane 0), per role."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
from pytorch_wavelets_b200 import _ffi

lib = _ffi.lib()
buf = (ctypes.c_ulonglong * 64)()
x = torch.randn(128, 32, 512, 512, device='cuda')
J = int(sys.argv[1]) if len(sys.argv) > 1 else 3
f = pw.DWTForward(J=J, wave='db4', mode='symmetric').cuda()
with torch.no_grad():
    f(x); f(x)
    lib.b200w_debug_pyr_prof(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(x); e1.record(); torch.cuda.synchronize()
    assert lib.b200w_debug_pyr_prof(buf) == 0
ms = e0.elapsed_time(e1)
names = {0: ['wait_empty', '', '', '', '', '', '', 'life'],
         1: ['flush', 'events', 'wait_read0', 'wait_final', '', '', '', 'life']}
out = {'ms': ms}
for role in range(2 + J):
    v = [buf[8 * role + k] for k in range(8)]
    nm = names.get(role, ['wait_in', 'wait_out_empty', 'wait_next_empty', 'fence', 'patch', 'colpass_emit', '', 'life'])
    life = max(v[7], 1)
    out['role%d' % role] = {n: round(v[k] / life, 4) if n not in ('events', 'life') else v[k] for k, n in enumerate(nm) if n}
print(json.dumps(out, indent=1))
