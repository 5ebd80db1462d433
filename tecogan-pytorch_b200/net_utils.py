"""Drop-in for the functional ops of codes/utils/net_utils.py (reference lines 36-156):
space_to_depth, backward_warp, get_upsampling_func, BicubicUpsampler -- same names, argument
meaning and error behaviour, executed by libtecogan_b200 on NCHW fp32 CUDA tensors.
"""
import torch
import torch.nn as nn

from . import lib as L
from . import ops


def needs_grad(*tensors):
    """True when autograd would have to differentiate through an op fed with these tensors."""
    return torch.is_grad_enabled() and any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def no_autograd(name, *tensors):
    """Ops without a backward kernel refuse inputs that require grad instead of silently cutting
    the graph (the reference's callers differentiate through backward_warp / upsample_func / fnet:
    vsr_model.py:86, vsrgan_model.py:106-108,214-222, tecogan_nets.py:419-420)."""
    if needs_grad(*tensors):
        raise NotImplementedError(
            f'tecogan-b200 {name}: no backward kernel for this op -- an input requires grad under '
            f'autograd; call it under torch.no_grad() or detach the input explicitly')


def _chk(t, name):
    if not t.is_cuda:
        raise L.TecoganB200Error(f'{name} must be a CUDA tensor: tecogan-b200 has no CPU path')
    return t


def _f32(t, name):
    if not t.is_cuda:
        raise L.TecoganB200Error(f'{name} must be a CUDA tensor: tecogan-b200 has no CPU path')
    return t.detach().float().contiguous()


def space_to_depth(x, scale):
    """Equivalent to tf.space_to_depth(): out[n,(sy*s+sx)*C+c,oh,ow] = x[n,c,oh*s+sy,ow*s+sx]."""
    if needs_grad(x):
        from .autograd import SpaceToDepthFunction
        return SpaceToDepthFunction.apply(_chk(x, 'x'), scale)
    return ops.space_to_depth(_f32(x, 'x'), scale)


def backward_warp(x, flow, mode='bilinear', padding_mode='border'):
    """Backward warp `x` (nchw) according to `flow` (n2hw): bilinear, border clamp,
    align_corners=True -- the only combination the reference ever uses."""
    if mode != 'bilinear' or padding_mode != 'border':
        raise ValueError(f'Unsupported warp mode: {mode}/{padding_mode}')
    if needs_grad(x, flow):
        # warp loss (vsr_model.py:86) and the ST-discriminator's input builder (tecogan_nets.py:452) train
        # through this op: scatter/gather backward kernels (tg_backward_warp_bwd_nchw_f32)
        from .autograd import WarpFunction
        return WarpFunction.apply(_chk(x, 'x'), _chk(flow, 'flow'))
    return ops.backward_warp(_f32(x, 'x'), _f32(flow, 'flow'))


class BicubicUpsampler(nn.Module):
    """TF-style bicubic (a=-0.75, no half-pixel shift, replicate border); keeps the reference's
    `kernels` buffer [scale,4] so BD checkpoints load strictly."""

    def __init__(self, scale_factor, a=-0.75):
        super().__init__()
        cubic = torch.tensor([[0, a, -2 * a, a], [1, 0, -(a + 3), a + 2],
                              [0, -a, (2 * a + 3), -(a + 2)], [0, 0, a, -a]], dtype=torch.float32)
        ks = [cubic @ torch.tensor([1.0, t, t * t, t * t * t]) for t in
              (1.0 * d / scale_factor for d in range(scale_factor))]
        self.scale_factor = scale_factor
        self.register_buffer('kernels', torch.stack(ks))
        if a != -0.75:
            raise ValueError('tecogan-b200 BicubicUpsampler is built for a=-0.75')

    def forward(self, input):
        if needs_grad(input):
            from .autograd import UpsampleFunction
            return UpsampleFunction.apply(_chk(input, 'input'), self.scale_factor, L.UP_BICUBIC)
        return ops.upsample(_f32(input, 'input'), self.scale_factor, L.UP_BICUBIC)


class BilinearUpsampler:
    """F.interpolate(scale_factor=s, mode='bilinear', align_corners=False) for BI degradation.
    A plain callable (not a Module), like the reference's functools.partial: no state_dict keys."""

    def __init__(self, scale_factor):
        self.scale_factor = scale_factor

    def __call__(self, input):
        if needs_grad(input):
            from .autograd import UpsampleFunction
            return UpsampleFunction.apply(_chk(input, 'input'), self.scale_factor, L.UP_BILINEAR)
        return ops.upsample(_f32(input, 'input'), self.scale_factor, L.UP_BILINEAR)


def get_upsampling_func(scale=4, degradation='BI'):
    if degradation == 'BI':
        return BilinearUpsampler(scale)
    if degradation == 'BD':
        return BicubicUpsampler(scale_factor=scale)
    raise ValueError(f'Unrecognized degradation type: {degradation}')


def up_mode_of(upsample_func):
    return L.UP_BICUBIC if isinstance(upsample_func, BicubicUpsampler) else L.UP_BILINEAR
