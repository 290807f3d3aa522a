"""world_size-2 gloo tests of the gradient exchange used by bench.py / training at N > 1 GPUs (stylegan_v_b200/optim.py::FlatModuleState)."""
import os
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker_flat_state(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from stylegan_v_b200.optim import FlatModuleState
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    conv = torch.nn.Conv2d(2, 2, 3)
    ema = [torch.nn.Parameter(p.detach().clone()) for p in list(lin.parameters()) + list(conv.parameters())]
    params = list(lin.parameters()) + list(conv.parameters())
    before = [p.detach().clone() for p in params]
    st = FlatModuleState(params, ema)
    same_values = all(torch.equal(p.detach(), b) for p, b in zip(params, before))          # re-homing keeps the values
    aligned = all(o % 64 == 0 for o in st.offsets) and params[1].data_ptr() == st.param.data_ptr() + st.offsets[1] * 4
    y = lin(torch.full((4, 5), float(rank + 1))).sum() + conv(torch.ones(1, 2, 5, 5) * (rank + 1)).sum()
    y.backward()
    local = st.grad.clone()
    st.all_reduce()                                               # SUM (the 1/world lives in the update kernel's grad_scale)
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = same_values and aligned and torch.allclose(st.grad, sum(gathered)) and st.world_size() == world
    # padding between parameters stays zero, so the flat update never sees garbage there
    mask = torch.ones_like(st.grad, dtype=torch.bool)
    for p, o in zip(params, st.offsets):
        mask[o:o + p.numel()] = False
    ok = ok and not st.grad[mask].any() and not st.param[mask].any()
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_flat_module_state_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_flat_state, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def _worker_early_bucket(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from stylegan_v_b200.optim import FlatModuleState
    torch.manual_seed(0)
    lin = torch.nn.Linear(5, 3)
    c1, c2 = torch.nn.Conv2d(2, 2, 3), torch.nn.Conv2d(2, 4, 3)
    params = list(lin.parameters()) + list(c1.parameters()) + list(c2.parameters())
    st = FlatModuleState(params, early=lambda p: p.ndim == 4)
    # the two conv weights lead the flat buffer; the bucket ends where the first other parameter starts
    ordered = st.params[0] is c1.weight and st.params[1] is c2.weight and st.early_numel == st.offsets[2] and 0 < st.early_numel < st.numel
    ok = ordered
    for it in range(2):                                               # two passes: the bucket re-arms
        st.zero_grad()
        x = torch.ones(1, 2, 7, 7) * (rank + 1 + it)
        loss = c2(c1(x)).sum() + lin(torch.full((4, 5), float(rank + 1))).sum()
        armed = st.begin_backward()
        loss.backward()
        fired_early = st._bucket['work'] is not None                  # issued from the hook of the last conv weight, before backward() returned
        local_rest = st.grad[st.early_numel:].clone()                 # (the early part may already hold the reduced values)
        st.finish_backward()
        # reference: plain local gradients summed over ranks
        st2_grads = torch.autograd.grad(c2(c1(x)).sum() + lin(torch.full((4, 5), float(rank + 1))).sum(), st.params)
        flat = torch.zeros_like(st.grad)
        for p, o, g in zip(st.params, st.offsets, st2_grads):
            flat[o:o + p.numel()] = g.reshape(-1)
        gathered = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        ok = ok and armed and fired_early and torch.allclose(st.grad, sum(gathered), rtol=1e-5, atol=1e-6) and torch.allclose(local_rest, flat[st.early_numel:])
    # without arming, finish_backward() is the single collective
    st.zero_grad()
    (lin(torch.ones(2, 5)).sum() * (rank + 1)).backward()
    st.finish_backward()
    ok = ok and torch.allclose(lin.bias.grad, torch.full((3,), 2.0 * (1 + 2)))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_early_gradient_bucket_world2():
    """FlatModuleState(early=...): the selected parameters lead the flat buffer and are all-reduced from the post-accumulate hook of the last
    of them (during backward); finish_backward() reduces the rest.  Result = one all-reduce of everything."""
    mgr = mp.Manager()
    out = mgr.dict()
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker_early_bucket, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]
