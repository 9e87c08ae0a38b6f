"""ScatLayer forward + backward timing at the configs[3] shape (training path; the BASELINE config itself is no-grad)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pytorch_wavelets_b200 as pw
x = torch.randn(256, 3, 256, 256, device='cuda', requires_grad=True)
s = torch.nn.Sequential(pw.ScatLayer(), pw.ScatLayer()).cuda()
def step():
    x.grad = None
    z = s(x)
    z.backward(torch.ones_like(z))
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step()
e1.record(); torch.cuda.synchronize()
print(json.dumps({'scat2_fwd_bwd_ms': round(e0.elapsed_time(e1) / 10, 4)}))
