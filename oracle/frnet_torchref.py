"""TEST / BENCH INFRASTRUCTURE ONLY -- the reference's CPU path restated with the SAME PyTorch
library operators the reference calls (F.conv2d, F.conv_transpose2d, F.max_pool2d,
F.interpolate, F.grid_sample, F.pad), functional over a state_dict.

Purpose: the timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`).  The closed-form
oracle in frnet_oracle.py is the parity checker (no dependence on grid_sample/interpolate
semantics) but its numpy sampling ops are slower than the library ops the reference really
runs; timing it would flatter the GPU.  tests/test_oracle_golden.py pins this module against
frnet_oracle and the reference-generated fixtures.

Reference operator sites: codes/models/networks/tecogan_nets.py:67-82 (FNet), :136-147 (SRNet),
:227-252 (step); codes/utils/net_utils.py:36-47, :50-82, :85-97, :133-156.
"""
import torch
import torch.nn.functional as F


def _c(x, p, k, act=None):
    y = F.conv2d(x, p[k + '.weight'], p[k + '.bias'], 1, 1)
    if act == 'lrelu':
        return F.leaky_relu(y, 0.2, inplace=True)
    if act == 'relu':
        return F.relu(y, inplace=True)
    return y


def upsample(p, x, scale, degradation):
    if degradation == 'BI':
        return F.interpolate(x, scale_factor=scale, mode='bilinear', align_corners=False)
    k = p['upsample_func.kernels']                       # [scale,4]
    n, c, h, w = x.shape
    z = F.pad(x.reshape(n * c, 1, h, w), (1, 2, 1, 2), mode='replicate')
    z = F.conv2d(z, k.view(scale, 1, 4, 1))              # vertical taps
    z = z.permute(0, 2, 1, 3).reshape(n * c, 1, scale * h, w + 3)
    z = F.conv2d(z, k.view(scale, 1, 1, 4))              # horizontal taps
    return z.permute(0, 2, 3, 1).reshape(n, c, scale * h, scale * w)


def fnet(p, x1, x2):
    o = torch.cat([x1, x2], 1)
    for e in ('encoder1', 'encoder2', 'encoder3'):
        o = F.max_pool2d(_c(_c(o, p, f'fnet.{e}.0', 'lrelu'), p, f'fnet.{e}.2', 'lrelu'), 2, 2)
    for d in ('decoder1', 'decoder2', 'decoder3'):
        o = F.interpolate(_c(_c(o, p, f'fnet.{d}.0', 'lrelu'), p, f'fnet.{d}.2', 'lrelu'),
                          scale_factor=2, mode='bilinear', align_corners=False)
    return torch.tanh(_c(_c(o, p, 'fnet.flow.0', 'lrelu'), p, 'fnet.flow.2')) * 24


def warp(x, flow):
    n, c, h, w = x.shape
    gx = torch.linspace(-1.0, 1.0, w, device=x.device, dtype=x.dtype).view(1, 1, 1, w).expand(n, -1, h, -1)
    gy = torch.linspace(-1.0, 1.0, h, device=x.device, dtype=x.dtype).view(1, 1, h, 1).expand(n, -1, -1, w)
    g = torch.cat([gx + flow[:, 0:1] / ((w - 1.0) / 2.0), gy + flow[:, 1:2] / ((h - 1.0) / 2.0)], 1)
    return F.grid_sample(x, g.permute(0, 2, 3, 1), mode='bilinear', padding_mode='border',
                         align_corners=True)


def s2d(x, s):
    n, c, h, w = x.shape
    return x.reshape(n, c, h // s, s, w // s, s).permute(0, 3, 5, 1, 2, 4).reshape(n, s * s * c, h // s, w // s)


def srnet(p, lr_curr, hr_tran, scale, degradation, nb):
    o = _c(torch.cat([lr_curr, hr_tran], 1), p, 'srnet.conv_in.0', 'relu')
    for i in range(nb):
        o = _c(_c(o, p, f'srnet.resblocks.{i}.conv.0', 'relu'), p, f'srnet.resblocks.{i}.conv.2') + o
    for u in range(2 if scale == 4 else 1):
        o = F.relu(F.conv_transpose2d(o, p[f'srnet.conv_up.{2 * u}.weight'], p[f'srnet.conv_up.{2 * u}.bias'],
                                      2, 1, output_padding=1), inplace=True)
    return _c(o, p, 'srnet.conv_out') + upsample(p, lr_curr, scale, degradation)


def step(p, lr_curr, lr_prev, hr_prev, scale, degradation, nb=10):
    flow = fnet(p, lr_curr, lr_prev)
    ph = lr_curr.size(2) - lr_curr.size(2) // 8 * 8
    pw = lr_curr.size(3) - lr_curr.size(3) // 8 * 8
    hr_flow = scale * upsample(p, F.pad(flow, (0, pw, 0, ph), 'reflect'), scale, degradation)
    return srnet(p, lr_curr, s2d(warp(hr_prev, hr_flow), scale), scale, degradation, nb)


def forward_sequence(p, lr_data, scale, degradation, nb=10):
    """FRNet.forward_sequence (tecogan_nets.py:174-225) with the reference's operators; fully
    differentiable through PyTorch autograd -- the CPU oracle of the generator BACKWARD
    (SURVEY.md 8-f1): gradients w.r.t. `p` and `lr_data` are pinned against reference-generated
    gradients in tests/test_oracle_golden.py."""
    n, t, c, h, w = lr_data.shape
    lr_prev = lr_data[:, :-1].reshape(n * (t - 1), c, h, w)
    lr_curr = lr_data[:, 1:].reshape(n * (t - 1), c, h, w)
    lr_flow = fnet(p, lr_curr, lr_prev)
    hr_flow = (scale * upsample(p, lr_flow, scale, degradation)).view(n, t - 1, 2, scale * h, scale * w)
    hr_prev = srnet(p, lr_data[:, 0], torch.zeros(n, scale * scale * c, h, w, dtype=lr_data.dtype,
                                                  device=lr_data.device), scale, degradation, nb)
    hr = [hr_prev]
    for i in range(1, t):
        hr_prev = srnet(p, lr_data[:, i], s2d(warp(hr_prev, hr_flow[:, i - 1]), scale), scale, degradation, nb)
        hr.append(hr_prev)
    return {'hr_data': torch.stack(hr, 1), 'hr_flow': hr_flow, 'lr_prev': lr_prev, 'lr_curr': lr_curr,
            'lr_flow': lr_flow}


def sequence_loss_and_grads(p, lr_data, scale, degradation, seed, nb=10):
    """loss = <hr_data, R1> + 0.05 <lr_flow, R2> with seeded uniform(-1,1) R1, R2; returns the loss
    and the gradients w.r.t. every floating-point parameter and lr_data (the protocol of the
    `fwd_seq_grads_*` fixture, oracle/gen_golden.py)."""
    import numpy as np
    q = {k: v.clone().requires_grad_(v.is_floating_point() and not k.endswith('kernels')) for k, v in p.items()}
    x = lr_data.clone().requires_grad_(True)
    d = forward_sequence(q, x, scale, degradation, nb)
    rng = np.random.default_rng(seed)
    r1 = torch.from_numpy(rng.uniform(-1, 1, size=tuple(d['hr_data'].shape)).astype(np.float32))
    r2 = torch.from_numpy(rng.uniform(-1, 1, size=tuple(d['lr_flow'].shape)).astype(np.float32))
    loss = (d['hr_data'] * r1).sum() + 0.05 * (d['lr_flow'] * r2).sum()
    names = [k for k, v in q.items() if v.requires_grad]
    grads = torch.autograd.grad(loss, [q[k] for k in names] + [x])
    return loss.detach(), dict(zip(names, grads[:-1])), grads[-1]

