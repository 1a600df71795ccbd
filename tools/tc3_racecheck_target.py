"""compute-sanitizer target: the tensor-core learner kernels and the output-layer kernels on small shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apex_b200 import _capi
L = _capi.lib()
s = torch.cuda.current_stream().cuda_stream
for M in (300, 1500):
    A = torch.randn(M, 256, device="cuda"); B = torch.relu(torch.randn(M, 256, device="cuda")); W = torch.randn(256, 256, device="cuda") / 16
    b = torch.randn(256, device="cuda"); C = torch.empty(M, 256, device="cuda"); G = torch.zeros(256, 256, device="cuda")
    x = torch.randn(M, 50, device="cuda"); G1 = torch.zeros(256, 50, device="cuda"); W1 = torch.randn(256, 50, device="cuda")
    for p in (3, 1):
        assert L.apex_tc3_linear(A.data_ptr(), 256, M, 256, W.data_ptr(), 256, 1, b.data_ptr(), 1, None, 0, C.data_ptr(), 256, p, s) == 0
        assert L.apex_tc3_linear(A.data_ptr(), 256, M, 256, W.data_ptr(), 1, 256, None, 0, B.data_ptr(), 256, C.data_ptr(), 256, p, s) == 0
        assert L.apex_tc3_linear(x.data_ptr(), 50, M, 50, W1.data_ptr(), 50, 1, b.data_ptr(), 1, None, 0, C.data_ptr(), 256, p, s) == 0
        assert L.apex_tc3_outer(A.data_ptr(), 256, B.data_ptr(), 256, 256, M, G.data_ptr(), 256, 1, p, s) == 0
        assert L.apex_tc3_outer(A.data_ptr(), 256, x.data_ptr(), 50, 50, M, G1.data_ptr(), 50, 1, p, s) == 0
torch.cuda.synchronize()
print("ok")
