import sys, torch
sys.path.insert(0, ".")
from apex_b200 import _capi
L = _capi.lib(); dev = torch.device("cuda:0")
rows, H = 65536, 256
x, w, b, y = torch.randn(rows, H, device=dev), torch.randn(H, H, device=dev) * 0.06, torch.randn(H, device=dev), torch.zeros(rows, H, device=dev)
for _ in range(3):
    L.apex_tc_linear_forward(x.data_ptr(), rows, H, w.data_ptr(), b.data_ptr(), H, 1, y.data_ptr(), None)
torch.cuda.synchronize()
