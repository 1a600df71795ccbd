"""world_size-2 gloo tests of the host-side sharding logic (no GPU): env-id ranges are disjoint across ranks, the
flattened-gradient all-reduce + 1/world scaling reproduces the single-process gradient, and the advantage moments
all-reduce gives the global mean / unbiased std."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w = torch.randn(7, 5)
    x = torch.randn(16, 5)
    y = torch.randn(16, 7)
    xs, ys = x[rank::world], y[rank::world]
    # per-rank gradient of a mean-squared loss over the local shard, flattened (the layout ppo.py all-reduces)
    wl = w.clone().requires_grad_(True)
    (((xs @ wl.t()) - ys) ** 2).mean().backward()
    g = wl.grad.reshape(-1).clone()
    dist.all_reduce(g)
    g *= 1.0 / world
    wf = w.clone().requires_grad_(True)
    (((x @ wf.t()) - y) ** 2).mean().backward()
    ok_grad = torch.allclose(g, wf.grad.reshape(-1), atol=1e-6)
    # advantage moments (sum, sum of squares, count)
    adv = torch.randn(64, dtype=torch.float64)
    a = adv[rank::world]
    mom = torch.tensor([a.sum(), (a * a).sum(), float(a.numel())], dtype=torch.float64)
    dist.all_reduce(mom)
    mean = mom[0] / mom[2]
    var = (mom[1] - mom[2] * mean * mean) / (mom[2] - 1)
    ok_mom = abs(mean - adv.mean()) < 1e-12 and abs(var.sqrt() - adv.std()) < 1e-12
    # env-id sharding: rank r owns [r*N, (r+1)*N)
    N = 8
    ids = torch.arange(rank * N, (rank + 1) * N)
    allids = [torch.zeros(N, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allids, ids)
    ok_ids = len(set(torch.cat(allids).tolist())) == world * N
    out[rank] = bool(ok_grad and ok_mom and ok_ids)
    dist.destroy_process_group()


def test_two_rank_gradient_and_moment_allreduce():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _eval_worker(rank, world, port, out):
    """Two ranks share the four command trials the reference's own tool was run on (tests/golden/evaltools.npz)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from apex_b200 import evaluate
    from tests.oracle_util import OracleBatchedEnv
    from tests.test_oracle_cpu import G, _rowwise, _torch_ref_actor
    g = np.load(os.path.join(G, "evaltools.npz"))
    made = []

    def env_fn(n):
        made.append(n)
        return OracleBatchedEnv(n)
    data = evaluate.eval_commands_sharded(env_fn, _rowwise(_torch_ref_actor()), g["speed_schedule"], g["orient_schedule"],
                                          num_steps=int(g["num_steps"]), max_speed=3, min_speed=0)
    out[rank] = bool(made == [2] and data.shape == (4, 6) and np.abs(data - g["command_rows"]).max() < 1e-12
                     and evaluate.rank_slice(5, 1, 2) == (3, 5) and evaluate.rank_slice(1, 1, 2) == (1, 1))
    dist.destroy_process_group()


def test_two_rank_sharded_eval_commands():
    """Trials shard across ranks with no data-path collective; one exchange of result rows; same rows as the reference's tool."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_eval_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def _ars_worker(rank, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    # ARS.step's exchange, through the functions it calls (apex_b200/ars.py): directions sharded by rank, [local, 2] return tables
    # all-gathered in direction order; every rank draws the same index stream and slices its own share
    from apex_b200.ars import gather_direction_returns, shard_of
    deltas = 8
    gen = torch.Generator(device="cpu").manual_seed(10)
    idx_all = torch.randint(0, 1000, (deltas,), generator=gen, dtype=torch.int64)
    idx_loc = shard_of(idx_all, rank, 2)
    r = torch.stack([idx_loc.double() * 0.5, idx_loc.double() * 0.25], dim=1)  # returns that identify their direction
    full = gather_direction_returns(r)
    ok = bool(torch.equal(full[:, 0], idx_all.double() * 0.5)) and bool(torch.equal(full[:, 1], idx_all.double() * 0.25))
    out[rank] = ok
    dist.destroy_process_group()


def test_ars_return_table_gathers_in_direction_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = 29600 + os.getpid() % 1000
    ps = [ctx.Process(target=_ars_worker, args=(r, port, out)) for r in range(2)]
    for p in ps:
        p.start()
    for p in ps:
        p.join(timeout=120)
    assert all(p.exitcode == 0 for p in ps) and out[0] and out[1]
