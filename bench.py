#!/usr/bin/env python
"""bench.py -- HR frames/sec of the FRNet hot path at 4x BD, LR 3x134x320 -> HR 3x536x1280.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|eager-gpu]

Workload (BASELINE.json configs[1]): TecoGAN 4x BD inference, synthetic 3x134x320 clips, 4 clips
lock-stepped per GPU.  One "step" = one recurrent frame of all 4 clips on one GPU = 4 HR frames.
N > 1: launched by torchrun, one rank per GPU; clips shard across ranks with no data-path
collective (the recurrence keeps a clip on one device) -> weak scaling.

Prints ONE JSON line (rank 0):
  value      whole-job HR frames/s with inputs resident in HBM (CUDA-graph replay of the step,
             timed with CUDA events, max over ranks)
  e2e        the same metric through the reference-facing call FRNet.infer_sequence() with HOST
             buffers: per step the H2D copy of the LR frames and the D2H copy of the uint8 HR
             frames are inside the timed region
  roofline   the dominant kernel (conv_chain_kernel: SRNet conv_in + 10 residual blocks = 21 convs
             64->64 in one persistent tcgen05 launch) timed live with CUDA events against the measured
             tensor peak; `traffic` = its DRAM bytes from one ncu --set full capture
  roofline_conv_single  one residual conv 64->64 as its own launch (conv_tcgen05_kernel)
  roofline_warp*  the fused warp+space_to_depth+concat kernel against the HBM roofline
  cpu_baseline   the reference's CPU path (oracle/frnet_torchref.py: same PyTorch CPU library
             ops as the reference) on the box's host cores, bounded sample (rank 0, N=1)

--impl reference times ONLY that CPU path with the same metric/unit (rank 0 alone).
--impl eager-gpu is a second REFERENCE arm (context, not part of the contract, never what `value`
measures): the same port of the reference's operator sequence, executed by PyTorch's CUDA library
kernels on the same GPU instead of the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LR = (3, 134, 320)
SCALE = 4
CLIPS_PER_GPU = 4
FLOP_PER_FRAME = 94.438e9            # reference counter, SURVEY.md 8-d (FNet 10.511 + SRNet 83.927)
RES_CONV_FLOP_PER_PX = 2 * 9 * 64 * 64
WARP_BYTES_PER_FRAME_FP32 = 22983680
PUBLISHED_FPS_1080TI = 27.0            # resources/benchmark.png (GTX 1080 Ti, batch 1)


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        d = json.load(open(path))
        return {'hbm_gbs': d['hbm_gbs'], 'tflops_burst': d['bf16_tflops'],
                'tflops_sustained': d['bf16_tflops_sustained'], 'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'src': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                 '-lms', '20'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def make_params():
    from oracle import frnet_oracle as O
    return O.make_frnet_params(0, scale=SCALE, degradation='BD', gain=1.0)


def synthetic_clips(n, t, seed=0):
    """n smooth translating clips [n,t,3,134,320] -- synthetic data."""
    import torch
    from oracle import frnet_oracle as O
    base = O.make_clip(seed, min(t, 12), *LR)            # generate 12 frames, then ping-pong in time
    idx = [i % (2 * len(base) - 2) for i in range(t)]
    idx = [i if i < len(base) else 2 * len(base) - 2 - i for i in idx]
    one = base[idx]
    return torch.stack([torch.roll(one, shifts=17 * k, dims=-1) for k in range(n)])


# =============================================================================== reference arm
def cpu_reference_fps(steps, warmup, threads=None):
    """Reference CPU path, 1 clip-frame per step (a bounded sample of the 4-clip step)."""
    import torch
    from oracle import frnet_torchref as R
    # torchrun exports OMP_NUM_THREADS=1; the reference arm is meant to use the host cores this
    # process may run on (affinity mask, capped at 64: oversubscribing oneDNN's OpenMP pool stalls)
    if threads is None and torch.get_num_threads() == 1:
        try:
            threads = min(64, len(os.sched_getaffinity(0)))
        except Exception:
            threads = None
    if threads:
        torch.set_num_threads(threads)
    p = make_params()
    g = torch.Generator().manual_seed(0)
    lr_curr = torch.rand(1, *LR, generator=g)
    lr_prev = torch.rand(1, *LR, generator=g)
    hr_prev = torch.rand(1, LR[0], SCALE * LR[1], SCALE * LR[2], generator=g)
    with torch.no_grad():
        for _ in range(warmup):
            R.step(p, lr_curr, lr_prev, hr_prev, SCALE, 'BD')
        t0 = time.perf_counter()
        for _ in range(steps):
            hr_prev = R.step(p, lr_curr, lr_prev, hr_prev, SCALE, 'BD')
        dt = time.perf_counter() - t0
    return steps / dt, dt, torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    fps, dt, cores = cpu_reference_fps(args.steps, max(args.warmup, 1))
    sample = (f'{args.steps} steps x 1 clip-frame 3x134x320 -> 3x536x1280 (1 of the {CLIPS_PER_GPU} '
              f'lock-stepped clips per step), reference CPU ops via oracle/frnet_torchref.py, fp32')
    line = {
        'impl': 'reference', 'metric': 'hr_frames_per_sec_4xBD_3x134x320', 'value': fps, 'unit': 'frames/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'TecoGAN 4x BD inference, synthetic 3x134x320 -> 3x536x1280, CPU reference path',
                   'clips_per_step': 1},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_eager_gpu(args, rank):
    """Context comparator, NOT part of the contract: the reference's operator sequence
    (oracle/frnet_torchref.py = the F.conv2d / grid_sample / interpolate calls the reference makes)
    executed by PyTorch's CUDA library kernels (cuDNN) on the same B200, same 4-clip step, timed with
    CUDA events.  Answers "what does the stock reference get on this GPU" (BASELINE.md section 5)."""
    if rank != 0:
        return
    import torch
    from oracle import frnet_torchref as R
    dev = torch.device('cuda', 0)
    torch.backends.cudnn.benchmark = True                      # codes/main.py:216
    g = torch.Generator().manual_seed(0)
    n = CLIPS_PER_GPU
    base = [torch.rand(n, *LR, generator=g), torch.rand(n, *LR, generator=g),
            torch.rand(n, LR[0], SCALE * LR[1], SCALE * LR[2], generator=g)]
    out = {}
    for name, dtype, tf32, cl in (('fp32', torch.float32, False, False), ('tf32', torch.float32, True, False),
                                  ('fp16_channels_last', torch.float16, True, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        p = {k: v.to(dev, dtype) for k, v in make_params().items()}
        lr_curr, lr_prev, hr_prev = (t.to(dev, dtype) for t in base)
        if cl:
            lr_curr, lr_prev, hr_prev = (t.contiguous(memory_format=torch.channels_last)
                                         for t in (lr_curr, lr_prev, hr_prev))
        with torch.no_grad():
            for _ in range(max(args.warmup, 3)):
                hr_prev = R.step(p, lr_curr, lr_prev, hr_prev, SCALE, 'BD')
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                hr_prev = R.step(p, lr_curr, lr_prev, hr_prev, SCALE, 'BD')
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[name] = {'ms_per_step': ms, 'frames_per_s': n * 1e3 / ms}
    print(json.dumps({'impl': 'eager-gpu', 'metric': 'hr_frames_per_sec_4xBD_3x134x320', 'unit': 'frames/s',
                      'clips_per_step': n, 'steps': args.steps, 'device': torch.cuda.get_device_name(0),
                      'note': 'reference operator sequence on PyTorch CUDA library kernels (cuDNN), no uint8/H2D',
                      'results': out}), flush=True)


# =============================================================================== our arm
# dram__bytes_read.sum + dram__bytes_write.sum of one conv_chain_kernel launch (ncu --set full), or None
CHAIN_DRAM_TRAFFIC = 23.582976e6 + 2.131712e6     # the 21 layers' activations stay in the 126 MB L2
CHAIN_DRAM_TRAFFIC_SRC = 'profiles/ncu_chain_r1u.md'


def _time_graph(fn, nbuf, reps, torch):
    """Average device time of one fn(i) launch: `reps` launches over `nbuf` rotating buffer sets are
    captured in a CUDA graph (so the number is the kernel, not the Python launch rate) and the
    replay is timed with CUDA events; 3 untimed replays first."""
    for i in range(nbuf):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i % nbuf)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def time_kernels(dev, pk):
    """Live CUDA-event timing of the two roofline kernels on rotating buffers larger than L2."""
    import torch
    import tecogan_b200 as T
    ops = sys.modules['tecogan-pytorch_b200.ops']
    L = sys.modules['tecogan-pytorch_b200.lib']
    n, (c, h, w) = CLIPS_PER_GPU, LR
    out = {}
    reps = 60
    # ---- dominant kernel: SRNet residual-block conv 64->64 (+bias, ReLU), n=4 frames per launch
    wt = torch.randn(64, 64, 3, 3, device=dev) * 0.04
    pc = ops.PackedConv(wt, torch.zeros(64, device=dev), L.CONV_3X3, L.ACT_RELU)
    nbuf = 10                                   # 10 x (22 MB in + 22 MB out) = 440 MB > 126 MB L2
    xs = [torch.randn(n, h, w, 64, device=dev).half() for _ in range(nbuf)]
    ys = [torch.empty_like(x) for x in xs]
    t_conv = _time_graph(lambda i: pc(xs[i], y=ys[i]), nbuf, reps, torch)
    flops = RES_CONV_FLOP_PER_PX * n * h * w
    out['roofline'] = {
        'kernel': 'conv_tcgen05_kernel<conv3x3, halo> (SRNet resblock conv 64->64, 4 frames/launch)',
        'bound': 'tensor', 'achieved': flops / t_conv / 1e12, 'peak': pk['tflops_burst'], 'unit': 'TFLOP/s',
        'frac': flops / t_conv / 1e12 / pk['tflops_burst'],
        # dram__bytes_read.sum + dram__bytes_write.sum of this launch, one `ncu --set full` capture
        # (profiles/ncu_conv_r1q.md, launch 1): 22.09 MB read (the fp16 input once) + 0.006 MB
        # written inside the measured window (the 22 MB output stays in the 126 MB L2)
        'traffic': 22.09152e6 + 0.006144e6, 'traffic_src': 'profiles/ncu_conv_r1q.md',
        'us_per_launch': t_conv * 1e6, 'flop_per_launch': flops,
        'peak_src': pk['src'] + ' burst (kernel timed alone)',
        'how': f'{reps} launches in one CUDA graph, {nbuf} rotating in/out pairs (440 MB > L2), CUDA events'}
    out['roofline_conv_single'] = out['roofline']
    # ---- dominant kernel of the step: conv_in + 10 residual blocks as ONE persistent launch
    if ops.chain_enabled():
        nl = 21
        pcs = [ops.PackedConv(torch.randn(64, 64, 3, 3, device=dev) * 0.04, torch.zeros(64, device=dev), L.CONV_3X3,
                              L.ACT_RELU if (i == 0 or i % 2 == 1) else L.ACT_NONE) for i in range(nl)]
        specs = [(pcs[0], 0, 1, None)]
        for b in range(10):
            specs += [(pcs[1 + 2 * b], 1, 2, None), (pcs[2 + 2 * b], 2, 1, 1)]
        chain = ops.ConvChain(specs)
        nb3 = 3                                  # 3 x (22 MB in + 2 x 22 MB work) = 198 MB > 126 MB L2
        sets = [[xs[i], ys[i], torch.empty_like(xs[i])] for i in range(nb3)]
        creps = 12
        t_chain = _time_graph(lambda i: chain(sets[i]), nb3, creps, torch)
        cflops = flops * nl
        out['roofline'] = {
            'kernel': 'conv_chain_kernel (SRNet conv_in + 10 residual blocks = 21 convs 64->64 in one persistent '
                      'launch, 4 frames/launch)',
            'bound': 'tensor', 'achieved': cflops / t_chain / 1e12, 'peak': pk['tflops_burst'], 'unit': 'TFLOP/s',
            'frac': cflops / t_chain / 1e12 / pk['tflops_burst'],
            'traffic': CHAIN_DRAM_TRAFFIC, 'traffic_src': CHAIN_DRAM_TRAFFIC_SRC,
            'us_per_launch': t_chain * 1e6, 'us_per_layer': t_chain * 1e6 / nl, 'flop_per_launch': cflops,
            'peak_src': pk['src'] + ' burst (kernel timed alone)',
            'how': f'{creps} launches in one CUDA graph, {nb3} rotating buffer sets (198 MB > L2), CUDA events'}
        del sets
    del xs, ys
    # ---- fused warp + space_to_depth + concat, HR flow given (BASELINE.md byte formula)
    H, W = SCALE * h, SCALE * w
    nb2 = 6                                      # 6 x 4 frames x ~20 MB = 470 MB > L2
    hp = [torch.rand(n, c, H, W, device=dev) for _ in range(nb2)]
    fl = [(torch.rand(n, 2, H, W, device=dev) - 0.5) * 6 for _ in range(nb2)]
    lr = [torch.rand(n, c, h, w, device=dev) for _ in range(nb2)]
    oo = [torch.empty(n, h, w, 64, dtype=torch.float16, device=dev) for _ in range(nb2)]
    lf = [(torch.rand(n, 2, h // 8 * 8, w // 8 * 8, device=dev) - 0.5) * 2 for _ in range(nb2)]
    for variant in ('hrflow', 'lrflow'):
        if variant == 'hrflow':
            call = lambda i: ops.warp_s2d_concat_hrflow(hp[i], fl[i], lr[i], SCALE, out=oo[i])
        else:
            call = lambda i: ops.warp_s2d_concat_lrflow(hp[i], lf[i], lr[i], SCALE, L.UP_BICUBIC, out=oo[i])
        t = _time_graph(call, nb2, reps, torch)
        alg = WARP_BYTES_PER_FRAME_FP32 * n
        moved = n * (c * H * W * 4 + (2 * H * W * 4 if variant == 'hrflow' else 2 * (h // 8 * 8) * (w // 8 * 8) * 4)
                     + c * h * w * 4 + h * w * 64 * 2)
        out['roofline_warp' if variant == 'hrflow' else 'roofline_warp_fused_lrflow'] = {
            'kernel': f'warp_s2d_concat_kernel<4,{variant}> (4 frames/launch)', 'bound': 'hbm',
            'achieved': alg / t / 1e9, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': alg / t / 1e9 / pk['hbm_gbs'],
            # one `ncu --set full` capture of the LR-flow variant (profiles/ncu_warp_r1x.md)
            'traffic': (36.009728e6 + 0.312576e6) if variant == 'lrflow' else None,
            'us_per_launch': t * 1e6, 'algorithmic_bytes_per_launch': alg,
            'bytes_actually_moved_per_launch': moved, 'moved_gbs': moved / t / 1e9,
            'peak_src': pk['src'], 'how': f'{reps} launches in one CUDA graph, {nb2} rotating buffer sets > L2'}
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import tecogan_b200 as T
    ops = sys.modules['tecogan-pytorch_b200.ops']

    assert torch.cuda.is_available(), 'bench.py needs a GPU (no CPU fallback exists)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG', 'WARN')          # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    pk = peaks()

    net = T.FRNet(3, 3, 64, 10, 'BD', SCALE)
    net.load_state_dict(make_params(), strict=True)
    net = net.to(dev).eval()
    n, (c, h, w) = CLIPS_PER_GPU, LR
    K, Wm = args.steps, max(args.warmup, 3)

    # ---------------- device-resident throughput: graph replay of the recurrent step
    eng = T.ClipEngine(net, n, c, h, w, dev)
    clips = synthetic_clips(n, 8, seed=rank).to(dev)           # [n,8,c,h,w] resident in HBM
    frames = clips.transpose(0, 1).contiguous()

    def step(i):
        p = i & 1
        eng.lr[p].copy_(frames[i % frames.shape[0]])           # device->device, 2 MB
        eng.run_frame(p)

    eng.reset()
    for i in range(Wm):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.LAUNCH_COUNT
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(K):
        step(Wm + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * K / (ms_max * 1e-3)
    launches_per_step = eng.launches_per_step + 0               # kernels inside one graph replay
    gpu_launches = launches_per_step * K

    if args.profile_only:          # under ncu: only the step loop, no JSON line
        if rank == 0:
            sampler.stop()
        return
    # ---------------- end to end through FRNet.infer_sequence with host buffers
    t_e2e = max(K, 4)
    host_clips = synthetic_clips(n, t_e2e, seed=100 + rank).pin_memory()     # [n,T,c,h,w] pinned
    net.infer_sequence(host_clips[:, :4], dev)                                # warm-up
    net.infer_sequence(host_clips, dev)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    w0 = time.perf_counter()
    seq = net.infer_sequence(host_clips, dev)                                 # uint8 [n,T,H,W,c] on host
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    assert seq.shape == (n, t_e2e, SCALE * h, SCALE * w, c) and str(seq.dtype) == 'uint8'
    te = torch.tensor([w1 - w0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * n * t_e2e / float(te.item())
    clocks = sampler.stop() if rank == 0 else None     # sampled over both timed regions

    line = None
    if rank == 0:
        roof = time_kernels(dev, pk)
        cpu = None
        if world == 1:
            steps_cpu = 24
            fps, dt, cores = cpu_reference_fps(steps_cpu, 2)
            cpu = {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                   'sample': f'{steps_cpu} clip-frames 3x134x320 -> 3x536x1280 (fp32, PyTorch CPU library ops as the '
                             f'reference uses, oracle/frnet_torchref.py), {dt:.1f} s of CPU work'}
        line = {
            'metric': 'hr_frames_per_sec_4xBD_3x134x320', 'value': value, 'unit': 'frames/s', 'n_gpus': world,
            'steps': K, 'warmup': Wm, 'ms_per_step': ms_max / K, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': value / PUBLISHED_FPS_1080TI, 'dtype': 'f16', 'data': 'synthetic',
            'config': {
                'workload': 'TecoGAN 4x BD inference, synthetic 3x134x320 -> 3x536x1280, batch=4 lock-stepped '
                            'clips per B200 (BASELINE.json configs[1]); clips shard across GPUs, no collective',
                'clips_per_gpu': n, 'frames_per_step': n * world, 'weights': 'seeded random init (no checkpoint)',
                'l2': 'per-step working set ~1.3 GB of activations (conv_up output alone 351 MB) >> 126 MB L2; '
                      'no explicit flush', 'conv_impl': ops.default_conv_impl(),
                'baseline_note': 'vs_baseline = value / 27 FPS published for 1x GTX 1080 Ti, batch 1 '
                                 '(resources/benchmark.png); no B200 number is published'},
            'gflop_per_frame': FLOP_PER_FRAME / 1e9,
            'model_tflops': value * FLOP_PER_FRAME / 1e12 / world,
            'model_tensor_frac_of_sustained': value * FLOP_PER_FRAME / 1e12 / world / pk['tflops_sustained'],
            'e2e': {'value': e2e_val, 'unit': 'frames/s', 'h2d_bytes_per_step': n * c * h * w * 4,
                    'd2h_bytes_per_step': n * SCALE * h * SCALE * w * c, 'steps': t_e2e,
                    'api': 'FRNet.infer_sequence(lr_data[n,t,c,h,w] pinned host) -> uint8 ndarray [n,t,H,W,c]'},
            'gpu_launches': gpu_launches, 'launches_per_step': launches_per_step,
            'clocks': clocks,
        }
        line.update(roof)
        if cpu is not None:
            line['cpu_baseline'] = cpu
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'eager-gpu'])
    ap.add_argument('--profile-only', action='store_true',
                    help='run only the device-resident step loop (for ncu captures); prints nothing')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        return run_reference(args, rank)
    if args.impl == 'eager-gpu':
        return run_eager_gpu(args, rank)
    run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
