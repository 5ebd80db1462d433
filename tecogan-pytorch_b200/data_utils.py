"""Data-side helpers of the path's callers, drop-in for
codes/utils/data_utils.py:11-53: the BD degradation (Gaussian blur + subsample) that
BaseModel.prepare_training_data / prepare_inference_data apply to GT frames
(codes/models/base_model.py:70-75, 110-118) -- on the device, no CPU fallback."""
import numpy as np
import torch

from . import lib as L
from . import ops


def create_kernel(sigma, ksize=None):
    """Same tensor as the reference's create_kernel: [3,3,k,k] fp32, the normalised 2-D Gaussian
    (k = 1 + 2*int(3*sigma)) on the channel diagonal, zeros elsewhere (data_utils.py:11-27)."""
    if ksize is None:
        ksize = 1 + 2 * int(sigma * 3.0)
    n = np.arange(ksize, dtype=np.float64) - (ksize - 1.0) / 2.0
    g = np.exp(-0.5 * (n / float(sigma)) ** 2)          # scipy.signal.windows.gaussian(ksize, std=sigma)
    k2 = np.outer(g, g)
    k2 = k2 / k2.sum()
    kernel = np.zeros((3, 3, ksize, ksize), dtype=np.float32)
    for c in range(3):
        kernel[c, c] = k2
    return torch.from_numpy(kernel)


_KERNEL_CACHE = {}


def _device_kernel2d(kernel, channels, device):
    """Validate that `kernel` is the depthwise (block-diagonal, identical blocks) weight the
    reference builds and return its [k,k] block on `device` (cached per tensor version)."""
    key = (kernel.data_ptr(), kernel._version, str(device))
    hit = _KERNEL_CACHE.get(key)
    if hit is not None:
        return hit
    if kernel.dim() != 4 or kernel.shape[0] != kernel.shape[1] or kernel.shape[2] != kernel.shape[3]:
        raise L.TecoganB200Error(f'downsample_bd: kernel shape {tuple(kernel.shape)} is not [c,c,k,k]')
    if kernel.shape[0] != channels:
        raise L.TecoganB200Error(f'downsample_bd: kernel is for {kernel.shape[0]} channels, data has {channels}')
    kc = kernel.detach().float().cpu()
    block = kc[0, 0]
    for a in range(kc.shape[0]):
        for b in range(kc.shape[1]):
            want = block if a == b else torch.zeros_like(block)
            if not torch.equal(kc[a, b], want):
                raise L.TecoganB200Error('downsample_bd: only the depthwise kernel of create_kernel() is supported '
                                         '(identical blocks on the diagonal, zeros elsewhere)')
    if len(_KERNEL_CACHE) > 16:
        _KERNEL_CACHE.clear()
    out = _KERNEL_CACHE[key] = block.contiguous().to(device)
    return out


def downsample_bd(data, kernel, scale, pad_data):
    """data [n,c,H,W] fp32 in [0,1] on a CUDA device -> blurred + subsampled [n,c,h,w]
    (data_utils.py:30-53: reflect pad when `pad_data`, then F.conv2d(data, kernel, stride=scale))."""
    if not isinstance(data, torch.Tensor) or not data.is_cuda:
        raise L.TecoganB200Error('downsample_bd: expected a CUDA tensor (no CPU fallback exists)')
    data = data.float().contiguous()
    return ops.downsample_bd(data, _device_kernel2d(kernel, data.shape[1], data.device), int(scale), bool(pad_data))
