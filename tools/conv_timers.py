#!/usr/bin/env python
"""Per-role cycle breakdown of conv_tcgen05_kernel (GPU only; debugging / profiling aid).

Registers a device buffer with tg_debug_set_conv_timers, runs one layer configuration and prints
the average over CTAs of every timer, per tile.  Usage:  python tools/conv_timers.py
"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tecogan_b200 as T   # noqa: E402

ops = sys.modules['tecogan-pytorch_b200.ops']
L = sys.modules['tecogan-pytorch_b200.lib']
NAMES = ['prod_wait_empty', 'mma_wait_tempty', 'mma_wait_full', 'mma_issue', 'mma_total', 'epi_wait_store',
         'epi_wait_tfull', 'epi_compute', 'epi_store', 'epi_total', 'kernel', 'prologue', 'tiles']


def run(name, cin, cout_real, h, w, n, kind=L.CONV_3X3, epilogue=L.EPI_NHWC_F16, residual=False, a_mode=None):
    dev = 'cuda:0'
    wshape = (cout_real, cin, 3, 3) if kind == L.CONV_3X3 else (cin, cout_real, 3, 3)
    pc = ops.PackedConv(torch.randn(*wshape, device=dev) * 0.05, torch.zeros(cout_real, device=dev), kind,
                        L.ACT_RELU if epilogue == L.EPI_NHWC_F16 else L.ACT_NONE, epilogue)
    x = torch.randn(n, h, w, cin, device=dev).half()
    res = torch.randn(n, h, w, pc.cout, device=dev).half() if residual else None
    y = None
    if epilogue == L.EPI_OUT_NCHW_F32:
        y = torch.zeros(n, cout_real, h, w, device=dev)
    for _ in range(3):
        y = pc(x, y=y, residual=res, a_mode=a_mode)
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    lib = L.load()
    lib.tg_debug_set_conv_timers(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pc(x, y=y, residual=res, a_mode=a_mode)
    e1.record()
    torch.cuda.synchronize()
    lib.tg_debug_set_conv_timers(ctypes.c_void_p(0))
    t = buf.view(148, 16).cpu().double()
    active = t[:, 10] > 0
    t = t[active]
    tiles = t[:, 12].mean().item()
    out = {'layer': name, 'us': e0.elapsed_time(e1) * 1e3, 'ctas': int(active.sum()), 'tiles_per_cta': tiles}
    for i, nm in enumerate(NAMES[:-1]):
        out[nm] = round(t[:, i].mean().item())
    out['per_tile'] = {nm: round(out[nm] / max(tiles, 1)) for nm in NAMES[:10]}
    print(json.dumps(out))
    return out


CHAIN_NAMES = ['prod_wait_flags', 'prod_wait_empty', 'mma_total', 'mma_wait', 'epi_wait_tfull', 'epi_total',
               'kernel', 'tiles', 'mma_issue0', 'mma_look', 'mma_issue1', 'mma_boundary',
               'chk_iters', 'chk_fence', 'chk_hits', 'chk_total']


def run_chain(blocks=10, n=4, h=134, w=320):
    """per-role timers of conv_chain_kernel (slots: tg_chain_tcgen05.cu CT_*)"""
    dev = 'cuda:0'
    nl = 1 + 2 * blocks
    pcs = [ops.PackedConv(torch.randn(64, 64, 3, 3, device=dev) * 0.04, torch.zeros(64, device=dev), L.CONV_3X3,
                          L.ACT_RELU if (i == 0 or i % 2 == 1) else L.ACT_NONE) for i in range(nl)]
    specs = [(pcs[0], 0, 1, None)]
    for b in range(blocks):
        specs += [(pcs[1 + 2 * b], 1, 2, None), (pcs[2 + 2 * b], 2, 1, 1)]
    chain = ops.ConvChain(specs)
    x = torch.randn(n, h, w, 64, device=dev).half()
    bufs = [x, torch.empty_like(x), torch.empty_like(x)]
    for _ in range(3):
        chain(bufs)
    buf = torch.zeros(148 * 16 + 8 * 24 * 16, dtype=torch.int64, device=dev)
    lib = L.load()
    lib.tg_debug_set_conv_timers(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    chain(bufs)
    e1.record()
    torch.cuda.synchronize()
    lib.tg_debug_set_conv_timers(ctypes.c_void_p(0))
    trace = buf[148 * 16:].view(-1, 8).cpu()
    buf = buf[:148 * 16]
    t = buf.view(148, 16).cpu().double()
    t = t[t[:, 6] > 0]
    tiles = t[:, 7].mean().item()
    out = {'layer': f'chain {nl} layers n{n} {h}x{w}', 'us': e0.elapsed_time(e1) * 1e3, 'ctas': int(t.shape[0]),
           'tile_layers_per_cta': tiles}
    for i, nm in enumerate(CHAIN_NAMES):
        if nm != 'tiles':
            out[nm] = round(t[:, i].mean().item())
    out['per_tile'] = {nm: round(out[nm] / max(tiles, 1)) for nm in CHAIN_NAMES if nm != 'tiles'}
    print(json.dumps(out))
    # distribution over CTAs: the CTAs that never wait for flags are the critical path
    full = buf.view(148, 16).cpu().double()
    order = torch.argsort(full[:, 0])
    for tag, idx in (('least flag wait', order[:3].tolist()), ('most flag wait', order[-3:].tolist())):
        for c in idx:
            print(json.dumps({'cta': c, 'which': tag, **{nm: int(full[c, i].item()) for i, nm in enumerate(CHAIN_NAMES)}}))
    # event trace of CTA 0: per tile [_, acc ready, stores issued, published, deps verified, TMA issued, MMA start, _]
    nmy = int(full[0, 7].item()) // nl
    tr = trace[:nmy * nl].double()
    def stat(x):
        x = x[torch.isfinite(x)]
        return [int(v) for v in torch.quantile(x, torch.tensor([0.1, 0.5, 0.9], dtype=torch.double)).tolist()] if len(x) else None
    sel = torch.arange(nmy, nmy * (nl - 1))
    ev = {
        'epilogue: acc ready -> stores issued (E2-E1)': tr[sel, 2] - tr[sel, 1],
        'publish: stores issued -> flag released (E3-E2)': tr[sel, 3] - tr[sel, 2],
        'own tile published -> next-layer same tile verified (E4[q]-E3[q-n])': tr[sel, 4] - tr[sel - nmy, 3],
        'verified -> TMA issued (E5-E4)': tr[sel, 5] - tr[sel, 4],
        'TMA issued -> MMA start (E6-E5)': tr[sel, 6] - tr[sel, 5],
        'MMA start -> acc ready (E1-E6)': tr[sel, 1] - tr[sel, 6],
        'MMA start(q+1) - MMA start(q)': tr[sel + 1, 6] - tr[sel, 6],
        'lead: TMA issue(q) - MMA start(q-3)': tr[sel, 5] - tr[sel - 3, 6],
    }
    print(json.dumps({'trace_cta0 p10/p50/p90 cycles': {k: stat(v) for k, v in ev.items()}}))
    q = torch.quantile(full[:, :16], torch.tensor([0.0, 0.5, 1.0], dtype=torch.double), dim=0)
    print(json.dumps({'min/med/max': {nm: [int(q[j, i].item()) for j in range(3)] for i, nm in enumerate(CHAIN_NAMES)}}))
    return out


def run_chain_ablate(n=4, h=134, w=320, blocks=10, max_ctas=0):
    """TIMING build of the chain with parts switched off (TG_CHAIN_ABLATE bits: 1 no dependency waits, 2 no epilogue
    global loads/stores, 4 no flag publication, 8 no TMA loads, 16 one MMA per tile, 32 no bias reads) -- the
    results are wrong on purpose; only the per-tile period is read."""
    dev = 'cuda:0'
    nl = 1 + 2 * blocks
    pcs = [ops.PackedConv(torch.randn(64, 64, 3, 3, device=dev) * 0.04, torch.zeros(64, device=dev), L.CONV_3X3,
                          L.ACT_RELU if (i == 0 or i % 2 == 1) else L.ACT_NONE) for i in range(nl)]
    specs = [(pcs[0], 0, 1, None)]
    for b in range(blocks):
        specs += [(pcs[1 + 2 * b], 1, 2, None), (pcs[2 + 2 * b], 2, 1, 1)]
    chain = ops.ConvChain(specs)
    x = torch.randn(n, h, w, 64, device=dev).half()
    bufs = [x, torch.empty_like(x), torch.empty_like(x)]
    for _ in range(3):
        chain(bufs)
    lib = L.load()
    buf = torch.zeros(148 * 16 + 8 * 24 * 16, dtype=torch.int64, device=dev)
    combos = [int(c) for c in os.environ.get('TG_ABLATE_COMBOS', '0,1,5,2,3,7,39,15,16,23,31,63,8,32,64,128,192,194').split(',')]
    for fl in combos:
        os.environ['TG_CHAIN_ABLATE'] = str(fl)
        buf.zero_()
        lib.tg_debug_set_conv_timers(ctypes.c_void_p(buf.data_ptr()))
        us = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            chain(bufs)
            e1.record()
            torch.cuda.synchronize()
            us.append(e0.elapsed_time(e1) * 1e3)
        lib.tg_debug_set_conv_timers(ctypes.c_void_p(0))
        t = buf[:148 * 16].view(148, 16).cpu().double()
        t = t[t[:, 6] > 0]
        tiles = t[:, 7].max().item()
        out = {'ablate': fl, 'us': round(min(us), 1), 'tile_layers_max': tiles,
               'kernel_cycles_per_tile': round(t[:, 6].max().item() / max(tiles, 1)),
               'per_tile': {nm: round(t[:, i].mean().item() / max(t[:, 7].mean().item(), 1)) for i, nm in enumerate(CHAIN_NAMES) if nm != 'tiles'}}
        print(json.dumps(out), flush=True)
    os.environ.pop('TG_CHAIN_ABLATE', None)


TAIL_NAMES = ['mma_wait_full', 'mma_wait_tempty', 'mma_wait_hrfull', 'mma_wait_d2empty', 'mma_total', 'epiA_wait',
              'epiA_busy', 'epiB_wait', 'epiB_tmem', 'epiB_exchange', 'epiB_residual', 'epiB_store', 'epiB_total',
              'kernel', 'tiles', 'epiB_top']


def run_tail(flags='0', n=4, h=268, w=640, with_lr=True, with_u8=True, accumulate=False):
    """per-role cycles of tail_tcgen05_kernel (fused ConvT + conv_out + residual + uint8), per tile"""
    dev = 'cuda:0'
    os.environ['TG_TAIL_FLAGS'] = flags
    up = ops.PackedConv(torch.randn(64, 64, 3, 3, device=dev) * 0.05, torch.zeros(64, device=dev), L.CONVT_3X3_S2, L.ACT_RELU)
    oc = ops.PackedConv(torch.randn(3, 64, 3, 3, device=dev) * 0.05, torch.zeros(3, device=dev), L.CONV_3X3, L.ACT_NONE,
                        L.EPI_OUT_NCHW_F32)
    x = torch.randn(n, h, w, 64, device=dev).half()
    lr = torch.rand(n, 3, h // 2, w // 2, device=dev) if with_lr else None
    y = torch.zeros(n, 3, 2 * h, 2 * w, device=dev)
    u8 = torch.zeros(n, 2 * h, 2 * w, 3, dtype=torch.uint8, device=dev) if with_u8 else None
    call = lambda: ops.fused_tail(up, oc, x, lr, 4, L.UP_BICUBIC, y=y, y_u8=u8, accumulate=accumulate)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call()
    e1.record()
    torch.cuda.synchronize()
    us_plain = e0.elapsed_time(e1) * 1e3
    buf = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    lib = L.load()
    lib.tg_debug_set_conv_timers(ctypes.c_void_p(buf.data_ptr()))
    call()
    torch.cuda.synchronize()
    lib.tg_debug_set_conv_timers(ctypes.c_void_p(0))
    t = buf.view(148, 16).cpu().double()
    t = t[t[:, 13] > 0]
    tiles = t[:, 14].mean().item()
    out = {'flags': flags, 'lr': with_lr, 'u8': with_u8, 'accumulate': accumulate, 'us': round(us_plain, 1), 'tiles_per_cta': tiles,
           'per_tile': {nm: round(t[:, i].mean().item() / max(tiles, 1)) for i, nm in enumerate(TAIL_NAMES) if nm != 'tiles'}}
    print(json.dumps(out), flush=True)
    return out


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'tail':
        run_tail('0')                                            # residual + uint8 inside the kernel
        run_tail('0', with_u8=False)                             # residual inside
        run_tail('0', with_lr=False, with_u8=False)              # conv only
        run_tail('0', with_lr=False, with_u8=False, accumulate=True)   # accumulate onto a pre-written residual
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'chain':
        run_chain()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'chain-ablate':
        run_chain_ablate()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'pair':
        for flags in ('0', '16'):
            os.environ['TG_DBG_FLAGS'] = flags
            run(f'res halo flags={flags} (16: no pair interleave)', 64, 64, 134, 320, 4)
            run(f'res halo+res flags={flags}', 64, 64, 134, 320, 4, residual=True)
            run(f'convT flags={flags}', 64, 64, 268, 640, 4, kind=L.CONVT_3X3_S2)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'align':
        for flags in ('0', '4', '12'):
            os.environ['TG_DBG_FLAGS'] = flags
            run(f'res halo flags={flags} (4: no tap shift, 8: SBO=1024)', 64, 64, 134, 320, 4)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'tapn':
        for flags in ('0', '1', '2', '3'):
            os.environ['TG_DBG_FLAGS'] = flags
            run(f'conv_out flags={flags}', 64, 3, 536, 1280, 4, epilogue=L.EPI_OUT_NCHW_F32)
        sys.exit(0)
    run('res 64->64 halo n4', 64, 64, 134, 320, 4)
    run('res 64->64 halo+residual n4', 64, 64, 134, 320, 4, residual=True)
    run('convT 64->64 268x640 n4', 64, 64, 268, 640, 4, kind=L.CONVT_3X3_S2)
    run('conv_out 64->3 536x1280 n4', 64, 3, 536, 1280, 4, epilogue=L.EPI_OUT_NCHW_F32)
    run('fnet 256->256 16x40 n4', 256, 256, 16, 40, 4)
