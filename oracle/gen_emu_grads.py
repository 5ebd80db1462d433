"""TEST INFRASTRUCTURE -- writes tests/golden/fwd_seq_grads_bd4_16x16_nb2_g15_fp16emu.npz: the gradients
of the same protocol as the reference-generated `fwd_seq_grads_*` fixture (oracle/gen_golden.py `grads`),
but computed by the CPU PRECISION MODEL of the CUDA training path: the real orchestration
(tecogan-pytorch_b200/autograd.py) over tests/fake_ops.py with fp16 storage of weights, activations and
loss-scaled gradients and fp32 accumulation.  The GPU test compares the CUDA result with the reference
fixture (distance = the fp16 design's, a few percent on this random-projection loss) AND with this
model (distance = accumulation order only), so a kernel error cannot hide inside the design tolerance.

    python oracle/gen_emu_grads.py
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

GRAD_FULL = ('fnet.encoder1.0.weight', 'fnet.flow.2.weight', 'fnet.flow.2.bias', 'srnet.conv_in.0.weight',
             'srnet.resblocks.1.conv.2.bias', 'srnet.conv_up.2.bias', 'srnet.conv_out.weight', 'srnet.conv_out.bias')


class _Patch:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def main():
    import tecogan_b200 as T
    import fake_ops
    import synthetic
    P = 'tecogan-pytorch_b200.'
    autograd = importlib.import_module(P + 'autograd')
    fake_ops.install(_Patch(), sys.modules[P + 'ops'], sys.modules[P + 'networks'], sys.modules[P + 'net_utils'], autograd)
    net = T.FRNet(3, 3, 64, 2, 'BD', 4)
    net.load_state_dict(synthetic.make_frnet_params(15, nb=2, scale=4, degradation='BD', gain=1.5), strict=True)
    net.train()
    lr = torch.from_numpy(np.random.default_rng(9).uniform(0, 1, size=(1, 3, 3, 16, 16)).astype(np.float32))
    d = net(lr)
    rng = np.random.default_rng(16)
    r1 = torch.from_numpy(rng.uniform(-1, 1, size=tuple(d['hr_data'].shape)).astype(np.float32))
    r2 = torch.from_numpy(rng.uniform(-1, 1, size=tuple(d['lr_flow'].shape)).astype(np.float32))
    loss = (d['hr_data'] * r1).sum() + 0.05 * (d['lr_flow'] * r2).sum()
    loss.backward()
    named = dict(net.named_parameters())
    out = {'loss': np.float32(loss.item()), 'names': np.array(list(named)),
           'norms': np.array([float(v.grad.norm()) for v in named.values()], np.float64)}
    for k in GRAD_FULL:
        out['g:' + k] = named[k].grad.numpy()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'fwd_seq_grads_bd4_16x16_nb2_g15_fp16emu.npz'), **out)
    print('fp16 precision-model grads: loss', loss.item())


if __name__ == '__main__':
    main()
