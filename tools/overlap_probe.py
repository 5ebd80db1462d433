#!/usr/bin/env python
"""GPU experiment behind the FNet(t+1) || SRNet(t) overlap: how does the SRNet chain
scale with the number of CTAs it may use, how slow is FNet on the SMs that are left, and what do the two
cost when they run concurrently on two streams?  Prints one JSON line per measurement.

    python tools/overlap_probe.py
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tecogan_b200 as T   # noqa: E402
import synthetic           # noqa: E402

ops = sys.modules['tecogan-pytorch_b200.ops']
L = sys.modules['tecogan-pytorch_b200.lib']
DEV = 'cuda:0'
N, H, W = 4, 134, 320


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def fnet_forward(fnet, x1, x2, ctas):
    a = ops.pack_pair(x1, x2)
    for name, _, _ in fnet.ENC:
        a = fnet._conv(name, 0, L.ACT_LRELU02)(a, max_ctas=ctas)
        a = fnet._conv(name, 2, L.ACT_LRELU02)(a, max_ctas=ctas)
        a = ops.maxpool2x2(a)
    for name, _, _ in fnet.DEC:
        a = fnet._conv(name, 0, L.ACT_LRELU02)(a, max_ctas=ctas)
        a = fnet._conv(name, 2, L.ACT_LRELU02)(a, max_ctas=ctas)
        a = ops.upsample2x(a)
    a = fnet._conv('flow', 0, L.ACT_LRELU02)(a, max_ctas=ctas)
    return fnet._conv('flow', 2, L.ACT_NONE, L.EPI_FLOW_NCHW_F32)(a, max_ctas=ctas)


def main():
    net = T.FRNet(3, 3, 64, 10, 'BD', 4)
    net.load_state_dict(synthetic.make_frnet_params(0), strict=True)
    net = net.to(DEV).eval()
    x = torch.randn(N, H, W, 64, device=DEV).half()
    lr = torch.rand(N, 3, H, W, device=DEV)
    lr2 = torch.rand(N, 3, H, W, device=DEV)
    with torch.no_grad():
        net.srnet.run_nhwc(x, lr)                     # builds the chain object
        chain = net.srnet._chain
        bufs = [x, torch.empty_like(x), torch.empty_like(x)]
        for ctas in (0, 140, 132, 120, 112, 100, 90, 74):
            us = timed(lambda: chain(bufs, max_ctas=ctas))
            print(json.dumps({'what': 'chain alone', 'max_ctas': ctas or 148, 'us': round(us, 1)}), flush=True)
        for ctas in (0, 74, 48, 36, 28):
            us = timed(lambda: fnet_forward(net.fnet, lr, lr2, ctas))
            print(json.dumps({'what': 'fnet alone (14 convs + helpers)', 'max_ctas': ctas or 148, 'us': round(us, 1)}), flush=True)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for c_chain, c_fnet in ((100, 48), (112, 36), (120, 28), (90, 58), (148, 148)):
            def both():
                ev = torch.cuda.Event()
                ev.record()
                s1.wait_event(ev)
                s2.wait_event(ev)
                with torch.cuda.stream(s1):
                    chain(bufs, max_ctas=c_chain)
                with torch.cuda.stream(s2):
                    fnet_forward(net.fnet, lr, lr2, c_fnet)
                torch.cuda.current_stream().wait_stream(s1)
                torch.cuda.current_stream().wait_stream(s2)
            us = timed(both, reps=8)
            print(json.dumps({'what': 'chain || fnet on two streams', 'chain_ctas': c_chain, 'fnet_ctas': c_fnet,
                              'us': round(us, 1)}), flush=True)


if __name__ == '__main__':
    main()
