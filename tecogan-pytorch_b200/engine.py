"""Clip streaming engine: n lock-stepped clips through the FRNet recurrence with static device
buffers, one CUDA graph per ping-pong parity, pinned host staging and copy streams.

Replaces the per-frame host loop of FRNet.infer_sequence (reference tecogan_nets.py:269-281):
the reference does an H2D copy, ~60 library launches, a device sync, a D2H copy and a NumPy
uint8 conversion per frame; here a frame is one graph replay, the uint8/HWC conversion is a
kernel (tg_float_to_uint8_nhwc) and the copies overlap compute on side streams.
"""
import collections
import os
import weakref

import numpy as np
import torch

from . import ops


def _use_graph():
    return os.environ.get('TECOGAN_B200_GRAPH', '1') != '0'


class ClipEngine:
    def __init__(self, net, n, c, h, w, device, use_graph=None):
        # weak reference: the engine cache below must not keep a dropped net (and its graphs) alive
        self._net = weakref.ref(net)
        self.n, self.c, self.h, self.w = n, c, h, w
        self.device = torch.device(device)
        s = net.scale
        self.H, self.W = s * h, s * w
        dev = self.device
        with torch.cuda.device(dev):
            self.lr = [torch.zeros(n, c, h, w, device=dev) for _ in range(2)]
            self.hr = [torch.zeros(n, c, self.H, self.W, device=dev) for _ in range(2)]
            self.u8 = [torch.empty(n, self.H, self.W, c, dtype=torch.uint8, device=dev) for _ in range(2)]
            self.use_graph = _use_graph() if use_graph is None else use_graph
            self.graphs = [None, None]
            self.main = torch.cuda.Stream(device=dev)
            self.h2d = torch.cuda.Stream(device=dev)
            self.d2h = torch.cuda.Stream(device=dev)
            self.launches_per_step = None
            self.stage = None
            if self.use_graph:
                self._capture()

    @property
    def net(self):
        net = self._net()
        if net is None:
            raise ops.L.TecoganB200Error('ClipEngine: the FRNet it was built for has been deleted')
        return net

    def close(self):
        """Drop the CUDA graphs (and their private memory pool) and the static buffers."""
        self.graphs = [None, None]
        self.lr = self.hr = self.u8 = []
        self.stage = None

    # one frame: parity p consumes lr[p] (current), lr[p^1] (previous), hr[p^1] -> hr[p], u8[p]
    def _enqueue(self, p):
        # float32_to_uint8 + CHW->HWC happen inside the step (fused SRNet tail, or one extra kernel)
        self.net.step_into(self.lr[p], self.lr[p ^ 1], self.hr[p ^ 1], self.hr[p], out_u8=self.u8[p])

    def _capture(self):
        with torch.cuda.device(self.device):
            warm = torch.cuda.Stream(device=self.device)
            warm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(warm):
                for p in (0, 1):            # builds packed weights, sets kernel attributes
                    self._enqueue(p)
            torch.cuda.current_stream().wait_stream(warm)
            torch.cuda.synchronize(self.device)
            pool = None
            for p in (0, 1):
                g = torch.cuda.CUDAGraph()
                before = ops.LAUNCH_COUNT
                with torch.cuda.graph(g, pool=pool):
                    self._enqueue(p)
                self.launches_per_step = ops.LAUNCH_COUNT - before
                pool = g.pool()
                self.graphs[p] = g
            self.reset()

    def reset(self):
        """lr_prev = hr_prev = 0 (reference tecogan_nets.py:269-270)."""
        with torch.cuda.device(self.device):
            for t in self.lr + self.hr:
                t.zero_()

    def run_frame(self, p):
        """Enqueue frame with parity p on the current stream."""
        if self.graphs[p] is not None:
            self.graphs[p].replay()
        else:
            before = ops.LAUNCH_COUNT
            self._enqueue(p)
            self.launches_per_step = ops.LAUNCH_COUNT - before

    def run_clips(self, lr_host, out_host=None):
        """lr_host: pinned CPU tensor (or CUDA tensor) [n,t,c,h,w] fp32; returns a pinned uint8
        tensor [t,n,H,W,c].  Three streams: H2D copies land in a 2-deep device staging ring (so the
        copy of frame i+1 overlaps the compute of frame i -- lr[p] itself is still being read as
        lr_prev), the main stream moves staging -> lr[p] (device-to-device, ~2 us) and replays the
        step graph, and the D2H stream drains the uint8 frames; one synchronisation at the end."""
        t = lr_host.shape[1]
        if out_host is None:   # caching host allocator: cheap after the first call
            out_host = torch.empty((t, self.n, self.H, self.W, self.c), dtype=torch.uint8,
                                   pin_memory=True)
        with torch.cuda.device(self.device):
            if self.stage is None:
                self.stage = [torch.empty_like(self.lr[0]) for _ in range(2)]
            self.main.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.main):
                self.reset()
            staged = [None, None]          # H2D into stage[q] finished
            consumed = [None, None]        # stage[q] copied into lr[] (free for the next H2D)
            out_copied = [None, None]      # D2H of u8[p] finished
            self.h2d.wait_stream(self.main)
            for i in range(t):
                p = q = i & 1
                with torch.cuda.stream(self.h2d):
                    if consumed[q] is not None:
                        self.h2d.wait_event(consumed[q])
                    for k in range(self.n):
                        self.stage[q][k].copy_(lr_host[k, i], non_blocking=True)
                    staged[q] = torch.cuda.Event()
                    staged[q].record(self.h2d)
                with torch.cuda.stream(self.main):
                    self.main.wait_event(staged[q])
                    self.lr[p].copy_(self.stage[q], non_blocking=True)
                    consumed[q] = torch.cuda.Event()
                    consumed[q].record(self.main)
                    if out_copied[p] is not None:
                        self.main.wait_event(out_copied[p])    # u8[p] free to overwrite
                    self.run_frame(p)
                    done = torch.cuda.Event()
                    done.record(self.main)
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(done)
                    out_host[i].copy_(self.u8[p], non_blocking=True)
                    out_copied[p] = torch.cuda.Event()
                    out_copied[p].record(self.d2h)
            self.d2h.synchronize()
            self.main.synchronize()
        return out_host


# Engines are cached per net (weakly: deleting the net frees its engines) and per clip geometry,
# least-recently-used first; at most TECOGAN_B200_MAX_ENGINES (default 3) geometries per net stay
# captured -- each holds two CUDA graphs and their pool (hundreds of MB to a few GB at video sizes).
_ENGINES = weakref.WeakKeyDictionary()


def _max_engines():
    return max(1, int(os.environ.get('TECOGAN_B200_MAX_ENGINES', '3')))


def _param_signature(net):
    p = next(net.parameters())
    return (str(p.device), p.data_ptr())


def get_engine(net, n, c, h, w, device):
    per_net = _ENGINES.get(net)
    sig = _param_signature(net)
    if per_net is None or per_net['sig'] != sig:
        # first use, or the parameters moved (net.to(other device) / re-materialised): captured graphs
        # would read freed buffers -> drop every engine of this net
        if per_net is not None:
            for eng in per_net['lru'].values():
                eng.close()
        per_net = _ENGINES[net] = {'sig': sig, 'lru': collections.OrderedDict()}
    lru = per_net['lru']
    key = (n, c, h, w, str(device))
    eng = lru.get(key)
    if eng is None:
        while len(lru) >= _max_engines():
            lru.popitem(last=False)[1].close()
        eng = lru[key] = ClipEngine(net, n, c, h, w, device)
    else:
        lru.move_to_end(key)
        # parameters may have changed since capture: repack in place (graphs read the same buffers)
        net.refresh_packed_weights()
    return eng


def release_engines(net=None):
    """Free the cached engines of `net` (all nets when None)."""
    nets = [net] if net is not None else list(_ENGINES.keys())
    for k in nets:
        per_net = _ENGINES.pop(k, None)
        if per_net is not None:
            for eng in per_net['lru'].values():
                eng.close()


def infer_clips(net, lr_data, device):
    """lr_data [n,t,c,h,w] fp32 (CPU or CUDA) -> np.uint8 [n,t,H,W,c]."""
    if lr_data.dim() != 5:
        raise ValueError('infer_clips expects ntchw')
    n, t, c, h, w = lr_data.shape
    device = torch.device(device)
    if device.type != 'cuda':
        raise ops.L.TecoganB200Error('tecogan-b200 runs on CUDA devices only (no CPU path)')
    eng = get_engine(net, n, c, h, w, device)
    src = lr_data.detach()
    if src.dtype != torch.float32:
        src = src.float()
    if not src.is_cuda and not src.is_pinned():
        src = src.pin_memory()           # pageable input: one staging copy (pass pinned to avoid)
    out = eng.run_clips(src)
    return out.numpy().transpose(1, 0, 2, 3, 4)
