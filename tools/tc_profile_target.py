"""ncu target: the hidden-layer GEMM of apex_mlp_forward_bf16 on the TMA route (k_tc_linear_tma), 65536 x 256 x 256."""
import sys, torch
sys.path.insert(0, ".")
from apex_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0")
rows, H = 65536, 256
p = lambda t: t.data_ptr()
x50, w1, b1 = torch.randn(rows, 50, device=dev), torch.randn(H, 50, device=dev) * 0.1, torch.randn(H, device=dev) * 0.1
w2, b2, w3, b3 = torch.randn(H, H, device=dev) * 0.06, torch.randn(H, device=dev), torch.randn(10, H, device=dev), torch.randn(10, device=dev)
h1, h2, y = torch.zeros(rows, H, device=dev), torch.zeros(rows, H, device=dev), torch.zeros(rows, 10, device=dev)
scratch = torch.zeros(L.apex_mlp_bf16_scratch_bytes(rows, H), dtype=torch.uint8, device=dev)
for _ in range(3):
    L.apex_mlp_forward_bf16(p(x50), rows, 50, H, 10, p(w1), p(b1), p(w2), p(b2), p(w3), p(b3), p(h1), p(h2), p(y), p(scratch), scratch.numel(), None)
torch.cuda.synchronize()
