"""World-size-2 gloo test (CPU) of the data-parallel exchange step (tecogan-pytorch_b200/reducer.py):
one flat all-reduce must give every rank the mean gradient and the mean of the logged scalars -- the same
result DistributedDataParallel + base_model.reduce_log produce -- and the clip sharding of the inference
path must partition the clips."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import tecogan_b200 as T  # noqa: F401
    red_mod = sys.modules['tecogan-pytorch_b200.reducer']
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, 1, 1), torch.nn.ReLU(), torch.nn.Conv2d(8, 3, 3, 1, 1))
    ref = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, 1, 1), torch.nn.ReLU(), torch.nn.Conv2d(8, 3, 3, 1, 1))
    ref.load_state_dict(net.state_dict())
    ddp = torch.nn.parallel.DistributedDataParallel(ref)
    red = red_mod.FlatGradientReducer(net, n_scalars=4)
    x = torch.randn(2, 3, 8, 8, generator=torch.Generator().manual_seed(10 + rank))
    ok = True
    for it in range(2):
        red.zero_grad()
        loss = net(x).square().mean() * (1 + it)
        loss.backward()
        logs = red.all_reduce_async({'l_pix_G': loss, 'it': float(it)}).wait()
        ddp.zero_grad()
        lref = ddp(x).square().mean() * (1 + it)
        lref.backward()
        ltot = lref.detach().clone()
        dist.all_reduce(ltot)
        for p, r in zip(net.parameters(), ref.parameters()):
            ok &= torch.allclose(p.grad, r.grad, rtol=1e-5, atol=1e-7)
            ok &= p.grad.data_ptr() >= red.flat.data_ptr()                      # still a view of the flat buffer
        ok &= abs(logs['l_pix_G'] - float(ltot) / world) < 1e-6 and abs(logs['it'] - it) < 1e-6
    q.put((rank, bool(ok), T.clips_for_rank(5, rank, world)))
    dist.destroy_process_group()


def test_flat_gradient_reducer_matches_ddp_world_size_2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(60)
    assert all(ok for _, ok, _ in res), res
    clips = res[0][2] + res[1][2]
    assert sorted(clips) == list(range(5)) and not set(res[0][2]) & set(res[1][2])


def test_single_process_reducer_is_a_no_op_exchange():
    import tecogan_b200 as T  # noqa: F401
    red_mod = sys.modules['tecogan-pytorch_b200.reducer']
    net = torch.nn.Linear(4, 2)
    red = red_mod.FlatGradientReducer(net, n_scalars=2)
    net(torch.ones(1, 4)).sum().backward()
    logs = red.all_reduce_async({'a': torch.tensor(3.0)}).wait()
    assert logs == {'a': 3.0} and torch.equal(net.weight.grad, torch.ones(2, 4))
    red.zero_grad()
    assert float(red.flat.abs().sum()) == 0.0 and net.weight.grad.data_ptr() == red.flat.data_ptr()
