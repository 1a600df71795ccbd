import sys, torch
sys.path.insert(0, ".")
from apex_b200.envs import BatchedCassieEnv
n = 2048
e64 = BatchedCassieEnv(n, dtype=torch.float64, seed=5, dynamics_randomization=True)
e32 = BatchedCassieEnv(n, dtype=torch.float32, seed=5, dynamics_randomization=True)
e64.reset(); e32.reset()
g = torch.Generator().manual_seed(0)
for k in range(12):
    act = torch.randn((n, 10), generator=g) * 0.2
    e32.st.copy_(e64.st.to(torch.float32)); e32.sti.copy_(e64.sti)
    e64.step(act.to("cuda", torch.float64)); e32.step(act.to("cuda"))
    c64, c32 = e64.field("cost").float().mean().item(), e32.field("cost").float().mean().item()
    i64, i32 = e64.field("solver_iter").float().mean().item(), e32.field("solver_iter").float().mean().item()
    print(f"step {k}: cost(sum iters*nefc over 50 substeps) f64 {c64:.0f} f32 {c32:.0f} ratio {c32/c64:.3f}; last-substep iters f64 {i64:.1f} f32 {i32:.1f}")
