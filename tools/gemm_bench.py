"""Hidden-layer GEMM y = relu(x W^T + b), 256 x 256: float32 SIMT kernels vs the tcgen05 bf16 kernel (CUDA events, L2 flushed by size)."""
import sys

import torch

sys.path.insert(0, ".")
from apex_b200 import _capi

L = _capi.lib()
dev = torch.device("cuda:0")
for rows in (4096, 65536):
    H = 256
    x, w, b = torch.randn(rows, H, device=dev), torch.randn(H, H, device=dev) * 0.06, torch.randn(H, device=dev)
    h1, y, y2 = torch.zeros(rows, H, device=dev), torch.zeros(rows, H, device=dev), torch.zeros(rows, 10, device=dev)
    w1, b1, w3, b3 = torch.randn(H, 50, device=dev), b, torch.randn(10, H, device=dev), torch.randn(10, device=dev)
    x50 = torch.randn(rows, 50, device=dev)
    p = lambda t: t.data_ptr()

    def timeit(fn, n=30):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    t_tc = timeit(lambda: L.apex_tc_linear_forward(p(x), rows, H, p(w), p(b), H, 1, p(y), None))
    t_f32 = timeit(lambda: L.apex_mlp_forward(p(x50), rows, 50, H, 10, p(w1), p(b1), p(w), p(b), p(w3), p(b3), p(h1), p(y), p(y2), None))
    t_bf = timeit(lambda: L.apex_mlp_forward_bf16(p(x50), rows, 50, H, 10, p(w1), p(b1), p(w), p(b), p(w3), p(b3), p(h1), p(y), p(y2), None, 0, None))
    scratch = torch.zeros(L.apex_mlp_bf16_scratch_bytes(rows, H), dtype=torch.uint8, device=dev)
    t_tma = timeit(lambda: L.apex_mlp_forward_bf16(p(x50), rows, 50, H, 10, p(w1), p(b1), p(w), p(b), p(w3), p(b3), p(h1), p(y), p(y2), p(scratch), scratch.numel(), None))
    wt = scratch[-H * H * 2:]
    t_layer = timeit(lambda: L.apex_tc_linear_tiled(p(scratch), rows, H, p(w), p(wt), p(b), H, 1, p(y), None))
    fl = 2.0 * rows * H * H
    print(f"rows {rows}: tcgen05 hidden layer {t_tc:.1f} us = {fl / t_tc / 1e6:.1f} TFLOP/s; whole MLP forward float32 {t_f32:.1f} us, "
          f"with tcgen05 hidden layer (convert on the fly) {t_bf:.1f} us, (TMA from tiled bf16) {t_tma:.1f} us; "
          f"TMA hidden layer alone {t_layer:.1f} us = {fl / t_layer / 1e6:.1f} TFLOP/s")
