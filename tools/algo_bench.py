"""Throughput of the two other learners at BASELINE.json's sizes (bring-up tool, not the bench contract; see bench.py).

    python tools/algo_bench.py td3 [updates]      configs[2]: TD3, 1M-transition device replay ring, batch sweep 256 .. 65536
    python tools/algo_bench.py ars [iterations]   configs[4]: ARS, 512 directions x 2 signs x 16 rollouts (per-GPU share under torchrun)

Synthetic inputs as SURVEY.md §8d lists them: the replay ring is filled with U[-1, 1] rows of 112 floats, the ARS policy is the
reference's zero-initialised 50 -> 32 -> 10 linear actor with sigma = 0.0075, horizon 400.  Timed with CUDA events on torch's
current stream (the stream every launch of these classes uses), after warm-up.  One JSON line per run.
NOTE: written at the end of round 1 after the GPU budget was spent — first run is due in round 2.
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def td3(updates):
    from apex_b200.td3 import TD3, ReplayBuffer
    rb = ReplayBuffer(50, 10, max_size=1_000_000)
    g = torch.Generator(device=rb.device).manual_seed(0)
    rb.storage.copy_(torch.rand(rb.storage.shape, generator=g, device=rb.device) * 2 - 1)
    rb.storage[:, -1] = (rb.storage[:, -1] > 0.9).float()  # done flags
    rb.size, rb.ptr = rb.max_size, 0
    algo = TD3(50, 10, 1.0, a_lr=3e-4, c_lr=1e-3)
    out = {"config": "TD3 Cassie-v0 sizes, 1M x 112 f32 replay ring (448 MB) on device", "sweep": []}
    for batch in (256, 1024, 4096, 16384, 65536):
        algo.train(rb, 5, batch_size=batch, generator=g)  # warm-up, allocates the batch buffers
        n = max(10, min(updates, updates * 4096 // batch))
        l0 = algo.launches
        ms = timed(lambda: algo.train(rb, 1, batch_size=batch, generator=g), n)
        out["sweep"].append({"batch": batch, "updates": n, "ms_per_update": ms, "updates_per_s": 1e3 / ms, "samples_per_s": batch * 1e3 / ms,
                             "gather_GBps_algorithmic": batch * 112 * 4 * 2 / ms / 1e6, "launches_per_update": (algo.launches - l0) / n})
    print(json.dumps(out))


def ars(iterations):
    import torch.distributed as dist
    if "RANK" in os.environ and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        dist.init_process_group("nccl")
    from apex_b200.ars import ARS, Linear_Actor
    from apex_b200.envs import BatchedCassieEnv
    dev = f"cuda:{torch.cuda.current_device()}"
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    algo = ARS(lambda: Linear_Actor(50, 10, 32), lambda n: BatchedCassieEnv(n, device=dev, seed=1, env_id0=rank * n), deltas=512, rollouts=16,
               step_size=0.02, std=0.0075, seed=3)
    algo.step(traj_len=16)  # warm-up
    steps, e0, e1 = 0, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0.record()
    for _ in range(iterations):
        steps += algo.step(traj_len=400)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "ARS Cassie-v0, 512 directions x 2 x 16 rollouts, linear 50-32-10 policy, horizon 400", "n_gpus": world,
                          "envs_per_gpu": algo.env.num_envs, "iterations": iterations, "env_steps": steps, "ms_per_iteration": float(ms) / iterations,
                          "env_steps_per_s": steps / float(ms) * 1e3}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "td3"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (200 if which == "td3" else 3)
    {"td3": td3, "ars": ars}[which](n)
