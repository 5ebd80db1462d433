"""TEST INFRASTRUCTURE ONLY -- closed-form numpy restatements of the hot-path
index/sampling ops.  fp32 arithmetic, NCHW layout like the reference.

Every function cites the reference lines it restates (paths relative to
/root/reference).  These are the bit-exact targets for the index math of the
CUDA kernels (space_to_depth / pixel-shuffle interleave / uint8 quantisation)
and the fp32 targets for the sampling ops (warp, bicubic, bilinear).
"""
import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------
# codes/utils/net_utils.py:36-47  space_to_depth (TF channel ordering)
# ----------------------------------------------------------------------------
def space_to_depth(x, scale):
    """out[n,(sy*s+sx)*C+c,oh,ow] = x[n,c,oh*s+sy,ow*s+sx]  (pure permutation).

    Note this is NOT F.pixel_unshuffle ordering (c*s*s+sy*s+sx).
    """
    n, c, in_h, in_w = x.shape
    s = scale
    oh, ow = in_h // s, in_w // s
    out = np.empty((n, s * s * c, oh, ow), dtype=x.dtype)
    for sy in range(s):
        for sx in range(s):
            blk = (sy * s + sx) * c
            out[:, blk:blk + c] = x[:, :, sy:oh * s:s, sx:ow * s:s]
    return out


# ----------------------------------------------------------------------------
# codes/utils/net_utils.py:50-82  backward_warp
#   grid = linspace(-1,1) + flow/((L-1)/2); F.grid_sample(bilinear, border,
#   align_corners=True)
# ----------------------------------------------------------------------------
def backward_warp(x, flow, exact_reference_grid=True):
    """Bilinear sample of x at (X+u, Y+v) with border clamping.

    exact_reference_grid=True follows the reference's fp32 round trip through
    the normalised [-1,1] grid (net_utils.py:62-72) and grid_sample's
    un-normalisation ((g+1)/2*(L-1)); False uses the cancelled closed form
    ix = X + u, iy = Y + v, which differs by ~1e-4 px of fp32 rounding.
    """
    x = np.asarray(x, dtype=F32)
    flow = np.asarray(flow, dtype=F32)
    n, c, h, w = x.shape
    if exact_reference_grid:
        iu = np.linspace(-1.0, 1.0, w, dtype=np.float64).astype(F32)  # torch.linspace fp32
        iv = np.linspace(-1.0, 1.0, h, dtype=np.float64).astype(F32)
        gx = iu[None, None, :] + flow[:, 0] / F32((w - 1.0) / 2.0)
        gy = iv[None, :, None] + flow[:, 1] / F32((h - 1.0) / 2.0)
        ix = ((gx + F32(1)) / F32(2)) * F32(w - 1)
        iy = ((gy + F32(1)) / F32(2)) * F32(h - 1)
    else:
        ix = np.arange(w, dtype=F32)[None, None, :] + flow[:, 0]
        iy = np.arange(h, dtype=F32)[None, :, None] + flow[:, 1]
    ix = np.clip(ix, F32(0), F32(w - 1)).astype(F32)
    iy = np.clip(iy, F32(0), F32(h - 1)).astype(F32)
    x0 = np.floor(ix).astype(np.int64)
    y0 = np.floor(iy).astype(np.int64)
    x1 = np.minimum(x0 + 1, w - 1)
    y1 = np.minimum(y0 + 1, h - 1)
    fx = (ix - x0.astype(F32)).astype(F32)
    fy = (iy - y0.astype(F32)).astype(F32)
    out = np.empty_like(x)
    nn = np.arange(n)[:, None, None]
    for ch in range(c):
        p = x[:, ch]
        v00 = p[nn, y0, x0]
        v01 = p[nn, y0, x1]
        v10 = p[nn, y1, x0]
        v11 = p[nn, y1, x1]
        # grid_sample weights: nw=(1-fx)(1-fy), ne=fx(1-fy), sw=(1-fx)fy, se=fx*fy
        out[:, ch] = (v00 * (1 - fx) * (1 - fy) + v01 * fx * (1 - fy)
                      + v10 * (1 - fx) * fy + v11 * fx * fy)
    return out.astype(F32)


# ----------------------------------------------------------------------------
# codes/utils/net_utils.py:101-156  BicubicUpsampler (TF-style, a=-0.75)
# ----------------------------------------------------------------------------
def bicubic_kernels(scale, a=-0.75):
    """kernels[d] = cubic @ [1, t, t^2, t^3], t = d/scale  (net_utils.py:116-131).

    For scale=4: [0,1,0,0], [-0.10546875,0.87890625,0.26171875,-0.03515625],
    [-0.09375,0.59375,0.59375,-0.09375], [-0.03515625,0.26171875,0.87890625,-0.10546875]
    -- all exactly representable in fp16/fp32.
    """
    cubic = np.array([[0, a, -2 * a, a],
                      [1, 0, -(a + 3), a + 2],
                      [0, -a, (2 * a + 3), -(a + 2)],
                      [0, 0, a, -a]], dtype=F32)
    ks = []
    for d in range(scale):
        t = F32(1.0 * d / scale)
        ks.append(cubic @ np.array([1, t, t * t, t * t * t], dtype=F32))
    return np.stack(ks).astype(F32)  # (scale, 4)


def bicubic_upsample(x, scale):
    """HR index Y = s*y + d samples LR rows clamp(y-1), y, clamp(y+1), clamp(y+2)
    (replicate pad (1,2,1,2), net_utils.py:141) with kernels[d]; vertical pass
    first (:144-146), then horizontal (:149-151).  No half-pixel shift."""
    x = np.asarray(x, dtype=F32)
    n, c, h, w = x.shape
    s = scale
    k = bicubic_kernels(s)
    ry = np.clip(np.arange(h)[:, None] + np.arange(-1, 3)[None, :], 0, h - 1)  # (h,4)
    rx = np.clip(np.arange(w)[:, None] + np.arange(-1, 3)[None, :], 0, w - 1)  # (w,4)
    # vertical: v[n,c,y,d,x] = sum_i k[d,i] * x[n,c,ry[y,i],x]
    v = np.zeros((n, c, h, s, w), dtype=F32)
    for i in range(4):
        v += k[None, None, None, :, i, None] * x[:, :, ry[:, i], :][:, :, :, None, :]
    v = v.reshape(n, c, h * s, w)
    o = np.zeros((n, c, h * s, w, s), dtype=F32)
    for j in range(4):
        o += k[None, None, None, None, :, j] * v[:, :, :, rx[:, j]][..., None]
    return o.reshape(n, c, h * s, w * s).astype(F32)


# ----------------------------------------------------------------------------
# F.interpolate(scale_factor=s, mode='bilinear', align_corners=False)
#   used by codes/utils/net_utils.py:87-89 (BI upsample_func) and
#   codes/models/networks/tecogan_nets.py:74-79 (FNet x2)
# ----------------------------------------------------------------------------
def bilinear_upsample(x, scale):
    x = np.asarray(x, dtype=F32)
    n, c, h, w = x.shape
    s = scale

    def axis(L):
        dst = np.arange(L * s, dtype=F32)
        src = np.maximum((dst + F32(0.5)) / F32(s) - F32(0.5), F32(0)).astype(F32)
        i0 = np.floor(src).astype(np.int64)
        i0 = np.minimum(i0, L - 1)
        i1 = np.minimum(i0 + 1, L - 1)
        f = (src - i0.astype(F32)).astype(F32)
        return i0, i1, f

    y0, y1, fy = axis(h)
    x0, x1, fx = axis(w)
    top = x[:, :, y0, :] * (1 - fy)[None, None, :, None] + x[:, :, y1, :] * fy[None, None, :, None]
    out = top[:, :, :, x0] * (1 - fx) + top[:, :, :, x1] * fx
    return out.astype(F32)


# ----------------------------------------------------------------------------
# codes/models/networks/tecogan_nets.py:239-241  F.pad(flow,(0,pw,0,ph),'reflect')
# ----------------------------------------------------------------------------
def reflect_pad_flow(flow, pad_h, pad_w):
    """padded row h8+i = row h8-2-i ; same for columns."""
    flow = np.asarray(flow)
    return np.pad(flow, ((0, 0), (0, 0), (0, pad_h), (0, pad_w)), mode='reflect')


# ----------------------------------------------------------------------------
# nn.MaxPool2d(2,2) floor  (tecogan_nets.py:28,35,42)
# ----------------------------------------------------------------------------
def maxpool2x2(x):
    n, c, h, w = x.shape
    ho, wo = h // 2, w // 2
    v = x[:, :, :ho * 2, :wo * 2].reshape(n, c, ho, 2, wo, 2)
    return v.max(axis=(3, 5))


# ----------------------------------------------------------------------------
# nn.ConvTranspose2d(nf,nf,3,2,1,output_padding=1)  (tecogan_nets.py:119-126)
# restated as 4 parity sub-convolutions whose outputs interleave like a
# pixel-shuffle.  wt layout [Cin, Cout, kH, kW].
# ----------------------------------------------------------------------------
CONVT_PARITY_TAPS = {
    # (py, px): [(dy, dx, ky, kx), ...]   out[2y+py,2x+px] += in[y+dy,x+dx] @ wt[:,:,ky,kx]
    (0, 0): [(0, 0, 1, 1)],
    (0, 1): [(0, 0, 1, 2), (0, 1, 1, 0)],
    (1, 0): [(0, 0, 2, 1), (1, 0, 0, 1)],
    (1, 1): [(0, 0, 2, 2), (0, 1, 2, 0), (1, 0, 0, 2), (1, 1, 0, 0)],
}


def conv_transpose3x3s2_parity(x, wt, bias):
    x = np.asarray(x, dtype=F32)
    n, cin, h, w = x.shape
    cout = wt.shape[1]
    xp = np.zeros((n, cin, h + 1, w + 1), dtype=F32)
    xp[:, :, :h, :w] = x
    out = np.zeros((n, cout, 2 * h, 2 * w), dtype=F32)
    for (py, px), taps in CONVT_PARITY_TAPS.items():
        acc = np.zeros((n, cout, h, w), dtype=F32)
        for (dy, dx, ky, kx) in taps:
            acc += np.einsum('nihw,io->nohw', xp[:, :, dy:dy + h, dx:dx + w],
                             wt[:, :, ky, kx].astype(F32), optimize=True)
        out[:, :, py::2, px::2] = acc + bias[None, :, None, None]
    return out


# ----------------------------------------------------------------------------
# codes/utils/data_utils.py:80-87  float32_to_uint8
# ----------------------------------------------------------------------------
def float32_to_uint8(x):
    """uint8(clip(round_half_even(x*255), 0, 255)) -- np.round is half-to-even."""
    return np.uint8(np.clip(np.round(np.asarray(x, dtype=F32) * F32(255)), 0, 255))


# ----------------------------------------------------------------------------
# fused stage the CUDA path produces in one kernel: SRNet's conv_in input
#   cat([lr_curr, space_to_depth(backward_warp(hr_prev, hr_flow), s)], 1)
#   (tecogan_nets.py:141,247,250)
# ----------------------------------------------------------------------------
def create_kernel(sigma, ksize=None):
    """The 2-D Gaussian of codes/utils/data_utils.py:11-27 as one [k,k] fp32 array (the reference
    stacks it on the diagonal of a [3,3,k,k] conv weight): k = 1 + 2*int(3*sigma),
    g[i] = exp(-0.5*((i-(k-1)/2)/sigma)^2) (scipy.signal.windows.gaussian), outer(g,g)/sum."""
    if ksize is None:
        ksize = 1 + 2 * int(sigma * 3.0)
    n = np.arange(ksize, dtype=np.float64) - (ksize - 1.0) / 2.0
    g = np.exp(-0.5 * (n / float(sigma)) ** 2)
    k2 = np.outer(g, g)
    return (k2 / k2.sum()).astype(np.float32)


def downsample_bd(data, kernel2d, scale, pad_data):
    """codes/utils/data_utils.py:30-53: optional reflect pad by (k-1)//2 before / k-1-(k-1)//2 after
    (F.pad 'reflect'), then a depthwise valid correlation with stride `scale`
    (F.conv2d(data, block_diag(kernel), stride=scale)).  data [n,c,H,W] fp32 -> [n,c,h,w]."""
    data = np.asarray(data, dtype=np.float32)
    k = kernel2d.shape[0]
    if pad_data:
        pt = (k - 1) // 2
        pb = (k - 1) - pt
        data = np.pad(data, ((0, 0), (0, 0), (pt, pb), (pt, pb)), mode='reflect')
    H, W = data.shape[2:]
    oh, ow = (H - k) // scale + 1, (W - k) // scale + 1
    out = np.zeros(data.shape[:2] + (oh, ow), dtype=np.float64)
    for i in range(k):
        for j in range(k):
            out += np.float64(kernel2d[i, j]) * data[:, :, i:i + (oh - 1) * scale + 1:scale,
                                                     j:j + (ow - 1) * scale + 1:scale]
    return out.astype(np.float32)


def warp_s2d_concat(hr_prev, hr_flow, lr_curr, scale, exact_reference_grid=True):
    w = backward_warp(hr_prev, hr_flow, exact_reference_grid)
    return np.concatenate([np.asarray(lr_curr, dtype=F32), space_to_depth(w, scale)], axis=1)
