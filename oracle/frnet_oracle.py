"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference FRNet.

Functional (no nn.Module): parameters travel as an OrderedDict keyed exactly
like the reference ``FRNet.state_dict()``, so a reference
checkpoint is directly usable.  Dense contractions use torch's CPU conv
(``F.conv2d`` -- the same third-party arithmetic the reference calls through
``nn.Conv2d``); every sampling / index op uses the closed forms of
``ops_oracle`` (numpy) so the oracle does not depend on grid_sample /
interpolate / ConvTranspose2d semantics.

Reference: codes/models/networks/tecogan_nets.py (FNet :16-82, ResidualBlock
:85-100, SRNet :103-147, FRNet :150-281), codes/utils/net_utils.py:36-156,
codes/utils/data_utils.py:80-87.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import ops_oracle as K


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _conv3x3(x, p, name, act=None):
    """nn.Conv2d(cin, cout, 3, 1, 1, bias=True) [+ activation]."""
    y = F.conv2d(x, p[name + '.weight'], p[name + '.bias'], stride=1, padding=1)
    if act == 'lrelu':      # nn.LeakyReLU(0.2)  -- FNet only (tecogan_nets.py:23-65)
        y = torch.where(y >= 0, y, y * 0.2)
    elif act == 'relu':     # nn.ReLU            -- SRNet (tecogan_nets.py:94,113,121,126)
        y = torch.clamp_min(y, 0)
    return y


# ----------------------------------------------------------------------------
# upsample_func  (codes/utils/net_utils.py:85-97)
# ----------------------------------------------------------------------------
def upsample_func(x, scale, degradation):
    if degradation == 'BD':
        return _t(K.bicubic_upsample(x.numpy(), scale))
    if degradation == 'BI':
        return _t(K.bilinear_upsample(x.numpy(), scale))
    raise ValueError(f'Unrecognized degradation type: {degradation}')


# ----------------------------------------------------------------------------
# FNet.forward  (tecogan_nets.py:67-82)
# ----------------------------------------------------------------------------
def fnet_forward(p, x1, x2, prefix='fnet.'):
    """flow from x1 (curr) to x2 (prev); output (8*floor(h/8), 8*floor(w/8))."""
    out = torch.cat([x1, x2], dim=1)
    for enc in ('encoder1', 'encoder2', 'encoder3'):
        out = _conv3x3(out, p, f'{prefix}{enc}.0', 'lrelu')
        out = _conv3x3(out, p, f'{prefix}{enc}.2', 'lrelu')
        out = _t(K.maxpool2x2(out.numpy()))
    for dec in ('decoder1', 'decoder2', 'decoder3'):
        out = _conv3x3(out, p, f'{prefix}{dec}.0', 'lrelu')
        out = _conv3x3(out, p, f'{prefix}{dec}.2', 'lrelu')
        out = _t(K.bilinear_upsample(out.numpy(), 2))
    out = _conv3x3(out, p, f'{prefix}flow.0', 'lrelu')
    out = _conv3x3(out, p, f'{prefix}flow.2', None)
    return torch.tanh(out) * 24  # 24 is the max velocity (tecogan_nets.py:80)


# ----------------------------------------------------------------------------
# SRNet.forward  (tecogan_nets.py:136-147)
# ----------------------------------------------------------------------------
def srnet_forward(p, lr_curr, hr_prev_tran, scale, degradation, nb=None,
                  prefix='srnet.', taps=None):
    out = _conv3x3(torch.cat([lr_curr, hr_prev_tran], dim=1), p, f'{prefix}conv_in.0', 'relu')
    if taps is not None:
        taps['conv_in'] = out
    if nb is None:
        nb = len({k.split('.')[2] for k in p if k.startswith(f'{prefix}resblocks.')})
    for i in range(nb):
        y = _conv3x3(out, p, f'{prefix}resblocks.{i}.conv.0', 'relu')
        y = _conv3x3(y, p, f'{prefix}resblocks.{i}.conv.2', None)
        out = y + out
    if taps is not None:
        taps['resblocks'] = out
    n_up = 2 if scale == 4 else 1
    for u in range(n_up):
        key = f'{prefix}conv_up.{2 * u}'
        out = _t(K.conv_transpose3x3s2_parity(out.numpy(), p[key + '.weight'].numpy(),
                                              p[key + '.bias'].numpy()))
        out = torch.clamp_min(out, 0)
    if taps is not None:
        taps['conv_up'] = out
    out = _conv3x3(out, p, f'{prefix}conv_out', None)
    out = out + upsample_func(lr_curr, scale, degradation)
    return out


# ----------------------------------------------------------------------------
# FRNet.step  (tecogan_nets.py:227-252)
# ----------------------------------------------------------------------------
def frnet_step(p, lr_curr, lr_prev, hr_prev, scale, degradation, taps=None,
               exact_reference_grid=True):
    lr_flow = fnet_forward(p, lr_curr, lr_prev)
    pad_h = lr_curr.size(2) - lr_curr.size(2) // 8 * 8
    pad_w = lr_curr.size(3) - lr_curr.size(3) // 8 * 8
    lr_flow_pad = _t(K.reflect_pad_flow(lr_flow.numpy(), pad_h, pad_w))
    hr_flow = scale * upsample_func(lr_flow_pad, scale, degradation)
    cat = _t(K.warp_s2d_concat(hr_prev.numpy(), hr_flow.numpy(), lr_curr.numpy(), scale,
                               exact_reference_grid))
    c = lr_curr.size(1)
    if taps is not None:
        taps['lr_flow'] = lr_flow
        taps['hr_flow'] = hr_flow
        taps['srnet_in'] = cat
    return srnet_forward(p, cat[:, :c], cat[:, c:], scale, degradation, taps=taps)


# ----------------------------------------------------------------------------
# FRNet.infer_sequence  (tecogan_nets.py:254-281)
# ----------------------------------------------------------------------------
def frnet_infer_sequence(p, lr_data, scale, degradation, return_float=False):
    """lr_data [T,C,h,w] fp32 -> uint8 [T,H,W,C]; state starts at zeros."""
    tot, c, h, w = lr_data.shape
    s = scale
    lr_prev = torch.zeros(1, c, h, w)
    hr_prev = torch.zeros(1, c, s * h, s * w)
    seq, fseq = [], []
    for i in range(tot):
        lr_curr = lr_data[i:i + 1]
        hr_curr = frnet_step(p, lr_curr, lr_prev, hr_prev, scale, degradation)
        lr_prev, hr_prev = lr_curr, hr_curr
        fseq.append(hr_curr[0].numpy())
        seq.append(K.float32_to_uint8(hr_curr[0].numpy()))
    out = np.stack(seq).transpose(0, 2, 3, 1)  # thwc
    if return_float:
        return out, np.stack(fseq)
    return out


# ----------------------------------------------------------------------------
# FRNet.forward_sequence  (tecogan_nets.py:174-225)  -- training forward
# ----------------------------------------------------------------------------
def frnet_forward_sequence(p, lr_data, scale, degradation):
    n, t, c, lr_h, lr_w = lr_data.shape
    hr_h, hr_w = lr_h * scale, lr_w * scale
    lr_prev = lr_data[:, :-1].reshape(n * (t - 1), c, lr_h, lr_w)
    lr_curr = lr_data[:, 1:].reshape(n * (t - 1), c, lr_h, lr_w)
    lr_flow = fnet_forward(p, lr_curr, lr_prev)
    hr_flow = scale * upsample_func(lr_flow, scale, degradation)
    hr_flow = hr_flow.view(n, t - 1, 2, hr_h, hr_w)
    hr_data = []
    hr_prev = srnet_forward(p, lr_data[:, 0], torch.zeros(n, scale * scale * c, lr_h, lr_w),
                            scale, degradation)
    hr_data.append(hr_prev)
    for i in range(1, t):
        cat = _t(K.warp_s2d_concat(hr_prev.numpy(), hr_flow[:, i - 1].numpy(),
                                   lr_data[:, i].numpy(), scale))
        hr_curr = srnet_forward(p, cat[:, :c], cat[:, c:], scale, degradation)
        hr_data.append(hr_curr)
        hr_prev = hr_curr
    return {
        'hr_data': torch.stack(hr_data, dim=1),
        'hr_flow': hr_flow,
        'lr_prev': lr_prev,
        'lr_curr': lr_curr,
        'lr_flow': lr_flow,
    }


# seeded weight / clip generators live in the neutral top-level module `synthetic` (bench.py and the
# product smoke path must not import oracle/ for their inputs); re-exported for the tests
from synthetic import frnet_param_shapes, make_frnet_params, make_clip  # noqa: E402,F401
