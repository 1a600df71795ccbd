import sys, torch
sys.path.insert(0, ".")
from apex_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0")
rows, H = 65536, 256
p = lambda t: t.data_ptr()
x50, w1, b1 = torch.randn(rows, 50, device=dev), torch.randn(H, 50, device=dev), torch.randn(H, device=dev)
w2, b2, w3, b3 = torch.randn(H, H, device=dev) * 0.06, torch.randn(H, device=dev), torch.randn(10, H, device=dev), torch.randn(10, device=dev)
h1, h2, y = torch.zeros(rows, H, device=dev), torch.zeros(rows, H, device=dev), torch.zeros(rows, 10, device=dev)
for _ in range(3):
    L.apex_mlp_forward(p(x50), rows, 50, H, 10, p(w1), p(b1), p(w2), p(b2), p(w3), p(b3), p(h1), p(h2), p(y), None)
torch.cuda.synchronize()
