"""Host-buffer front end: run a transform module over a batch that lives in pinned HOST memory, with the
host->device copy of the inputs, the kernels and the device->host copy of every output overlapped on three
CUDA streams, chunked along N.  This is the end-to-end path bench.py times (`e2e`): PCIe-bound by design.
"""
import torch


def _flatten(out):
    if isinstance(out, torch.Tensor):
        return [out]
    flat = []
    for o in out:
        if o is None:
            continue
        flat.extend(_flatten(o))
    return flat


class HostPipeline(object):
    def __init__(self, module, shape, device, chunk=16):
        self.module = module
        self.device = torch.device(device)
        self.N = shape[0]
        self.chunk = min(chunk, self.N)
        self.s_h2d = torch.cuda.Stream(self.device)
        self.s_cmp = torch.cuda.Stream(self.device)
        self.s_d2h = torch.cuda.Stream(self.device)
        self.dev_in = [torch.empty((self.chunk,) + tuple(shape[1:]), device=self.device) for _ in range(2)]
        self.free_ev = [None, None]
        with torch.no_grad():
            probe = _flatten(module(self.dev_in[0]))
        self.host_out = []
        for t in probe:
            if t.dim() == 0:
                self.host_out.append(None)
                continue
            assert t.shape[0] == self.chunk, 'outputs must be batch-major (default o_dim / ri_dim)'
            self.host_out.append(torch.empty((self.N,) + tuple(t.shape[1:]), dtype=t.dtype).pin_memory())
        self.out_bytes = sum(t.numel() * t.element_size() for t in self.host_out if t is not None)
        del probe
        torch.cuda.synchronize(self.device)

    def run(self, host_x):
        """host_x: pinned (N, C, H, W) float32.  Returns the list of pinned host output tensors.  The copies are
        asynchronous: the current STREAM waits for them before this returns, the HOST does not -- call
        ``self.done.synchronize()`` (or ``torch.cuda.synchronize()``) before reading the host tensors on the CPU.
        The output buffers are reused by the next ``run``; it waits for the previous one's copies first."""
        if getattr(self, 'done', None) is not None:
            self.done.synchronize()
        cur = torch.cuda.current_stream(self.device)
        for s in (self.s_h2d, self.s_cmp, self.s_d2h):
            s.wait_stream(cur)
        keep = []
        for ci, n0 in enumerate(range(0, self.N, self.chunk)):
            n1 = min(n0 + self.chunk, self.N)
            buf = self.dev_in[ci % 2][:n1 - n0]
            with torch.cuda.stream(self.s_h2d):
                if self.free_ev[ci % 2] is not None:
                    self.s_h2d.wait_event(self.free_ev[ci % 2])
                buf.copy_(host_x[n0:n1], non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.s_h2d)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_in)
                with torch.no_grad():
                    outs = _flatten(self.module(buf))
                ev_done = torch.cuda.Event()
                ev_done.record(self.s_cmp)
                self.free_ev[ci % 2] = ev_done
            with torch.cuda.stream(self.s_d2h):
                self.s_d2h.wait_event(ev_done)
                for o, h in zip(outs, self.host_out):
                    if h is None:
                        continue
                    o.record_stream(self.s_d2h)
                    h[n0:n1].copy_(o, non_blocking=True)
            keep.append(outs)
        self.done = torch.cuda.Event()
        self.done.record(self.s_d2h)
        cur.wait_stream(self.s_d2h)
        cur.wait_stream(self.s_cmp)
        return self.host_out
