"""Host-buffer front end: chunked H2D -> synthesis -> D2H pipeline on three CUDA streams.

The synthesis kernels take ~1.7 ms for 32 x 10 s utterances while the PCIe copies of the controls
(70 MB in) and of the waveform (56 MB out) take longer than that, so a caller that starts from host
memory (the reference's main.py does: main.py:201-215 copies in, :272 copies out) is bound by
the copies unless they overlap.  ``HostPipeline`` splits the batch into chunks of utterances
(independent end to end) and overlaps the upload of chunk c+1, the kernels of chunk c and the
download of chunk c-1.  PCIe is full duplex, so the steady state costs max(H2D, kernels, D2H).

What is left outside the steady state is the upload of the FIRST chunk (nothing to compute yet) and the download
of the LAST one (nothing left to compute), while every extra chunk costs ~0.1 ms of partially filled kernel
waves.  ``chunks`` may therefore be a list of relative chunk sizes: a tapered schedule (small first and last
chunk, large middle ones, e.g. ``(4, 9, 13, 6)``) shortens fill and drain without adding launches.
"""
import torch


def chunk_bounds(batch, chunks):
    """[(lo, hi)] covering range(batch): ``chunks`` = number of (nearly) equal chunks, or a sequence of relative
    sizes (largest-remainder rounding; empty chunks are dropped)."""
    if isinstance(chunks, int):
        n = max(1, min(chunks, batch))
        base, rem = divmod(batch, n)
        sizes = [base + (1 if c < rem else 0) for c in range(n)]
    else:
        w = [float(x) for x in chunks]
        if not w or min(w) < 0 or sum(w) <= 0:
            raise ValueError("chunk weights must be non-negative and not all zero")
        exact = [x * batch / sum(w) for x in w]
        sizes = [int(e) for e in exact]
        for i in sorted(range(len(w)), key=lambda i: exact[i] - sizes[i], reverse=True)[:batch - sum(sizes)]:
            sizes[i] += 1
    bounds, lo = [], 0
    for n in sizes:
        if n > 0:
            bounds.append((lo, lo + n))
            lo += n
    assert lo == batch
    return bounds


class HostPipeline:
    def __init__(self, device, chunks=4, compute_streams=1):
        """``compute_streams`` > 1 (opt-in, not yet measured): consecutive chunks run on alternating side streams, so the
        small frame-rate kernels of chunk c+1 (which fill a fraction of the GPU) can overlap the FIR of chunk c instead
        of queueing behind it.  1 = everything on the caller's current stream (the measured configuration)."""
        self.device = torch.device(device)
        self.chunks = int(chunks) if isinstance(chunks, int) else tuple(chunks)
        self.h2d = torch.cuda.Stream(device=self.device)
        self.d2h = torch.cuda.Stream(device=self.device)
        self.compute = [torch.cuda.Stream(device=self.device) for _ in range(int(compute_streams))] \
            if int(compute_streams) > 1 else []
        self._dev = {}

    def _device_like(self, name, host):
        t = self._dev.get(name)
        if t is None or t.shape != host.shape or t.dtype != host.dtype:
            t = torch.empty(host.shape, dtype=host.dtype, device=self.device)
            self._dev[name] = t
        return t

    def run(self, host_inputs, forward_chunk, out_host):
        """host_inputs: dict name -> pinned host tensor with the batch in dim 0.
        forward_chunk(dev_inputs: dict of device views [lo:hi], lo, hi) -> device tensor [hi-lo, ...]
        out_host: pinned host tensor [B, ...] receiving the result.
        Returns an event recorded after the last download (call .synchronize() before reading out_host)."""
        for k, v in host_inputs.items():
            if not v.is_pinned():
                raise ValueError("host input %r must be in pinned memory for asynchronous copies" % k)
        if not out_host.is_pinned():
            raise ValueError("out_host must be in pinned memory")
        B = out_host.shape[0]
        main = torch.cuda.current_stream(self.device)
        dev = {k: self._device_like(k, v) for k, v in host_inputs.items()}
        bounds = chunk_bounds(B, self.chunks)
        self.h2d.wait_stream(main)          # device buffers may still be in use by earlier work on `main`
        self.d2h.wait_stream(main)
        up = []
        for lo, hi in bounds:               # all uploads are queued up front: they run back to back on the copy engine
            with torch.cuda.stream(self.h2d):
                for k, v in host_inputs.items():
                    dev[k][lo:hi].copy_(v[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.h2d)
            up.append(ev)
        for s in self.compute:
            s.wait_stream(main)
        for c, ((lo, hi), ev) in enumerate(zip(bounds, up)):
            cs = self.compute[c % len(self.compute)] if self.compute else main
            with torch.cuda.stream(cs):                 # a no-op context when cs is the current stream
                cs.wait_event(ev)
                out = forward_chunk({k: t[lo:hi] for k, t in dev.items()}, lo, hi)
                done = torch.cuda.Event()
                done.record(cs)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(done)
                out_host[lo:hi].copy_(out, non_blocking=True)
                out.record_stream(self.d2h)
        fin = torch.cuda.Event()
        fin.record(self.d2h)
        main.wait_stream(self.d2h)
        for s in self.compute:
            main.wait_stream(s)
        return fin
