"""ncu target: one MLP forward + backward at the PPO update's actor shape (65536 rows) on the tensor-core route."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apex_b200 import _capi

L = _capi.lib()
L.apex_set_tc_mode(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
rows, din, hid, dout = int(sys.argv[2]) if len(sys.argv) > 2 else 65536, 50, 256, 10
s = torch.cuda.current_stream().cuda_stream
x = torch.randn(rows, din, device="cuda")
w1, b1 = torch.randn(hid, din, device="cuda") / 7, torch.zeros(hid, device="cuda")
w2, b2 = torch.randn(hid, hid, device="cuda") / 16, torch.zeros(hid, device="cuda")
w3, b3 = torch.randn(dout, hid, device="cuda") / 16, torch.zeros(dout, device="cuda")
h1, h2, y = (torch.empty(rows, n, device="cuda") for n in (hid, hid, dout))
dy = torch.randn(rows, dout, device="cuda")
dh2, dh1 = torch.empty(rows, hid, device="cuda"), torch.empty(rows, hid, device="cuda")
gw1, gb1, gw2, gb2, gw3, gb3 = (torch.zeros_like(t) for t in (w1, b1, w2, b2, w3, b3))
for _ in range(3):
    L.apex_mlp_forward(x.data_ptr(), rows, din, hid, dout, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), w3.data_ptr(), b3.data_ptr(),
                       h1.data_ptr(), h2.data_ptr(), y.data_ptr(), s)
    L.apex_mlp_backward(x.data_ptr(), rows, din, hid, dout, w2.data_ptr(), w3.data_ptr(), h1.data_ptr(), h2.data_ptr(), dy.data_ptr(), dh2.data_ptr(),
                        dh1.data_ptr(), gw1.data_ptr(), gb1.data_ptr(), gw2.data_ptr(), gb2.data_ptr(), gw3.data_ptr(), gb3.data_ptr(), s)
torch.cuda.synchronize()
