import os, sys, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apex_b200 import _capi
L = _capi.lib()
L.apex_tc3_set_debug.argtypes = [ctypes.c_int]
s = lambda: torch.cuda.current_stream().cuda_stream
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
M = 65536
A = torch.randn(M, 256, device="cuda"); W = torch.randn(256, 256, device="cuda") / 16; b = torch.randn(256, device="cuda"); C = torch.empty(M, 256, device="cuda")
for p in (3, 1):
    for dbg in (0, 0, 2, 4, 6):
        L.apex_tc3_set_debug(dbg)
        t = timeit(lambda: L.apex_tc3_linear(A.data_ptr(), 256, M, 256, W.data_ptr(), 256, 1, b.data_ptr(), 1, None, 0, C.data_ptr(), 256, p, s()))
        print(f"passes {p} dbg {dbg} (1=noB 2=noA 4=noStore): {t:.1f} us", flush=True)
L.apex_tc3_set_debug(0)
B = torch.relu(torch.randn(M, 256, device="cuda")); G = torch.zeros(256, 256, device="cuda")
for p in (3, 1):
    for dbg in (0, 2, 4, 6):
        L.apex_tc3_set_debug(dbg)
        t = timeit(lambda: L.apex_tc3_outer(A.data_ptr(), 256, B.data_ptr(), 256, 256, M, G.data_ptr(), 256, 1, p, s()))
        print(f"outer passes {p} dbg {dbg} (2=noLoads 4=noAtomics): {t:.1f} us", flush=True)
L.apex_tc3_set_debug(0)
