"""The split-tf32 tensor-core GEMMs (csrc/tc_gemm3.cu) against float64: the 256-wide layers of the reference's actor / critic
(rl/policies/actor.py:142-215) computed on tcgen05 must stay float32-accurate (mode 3), and the plain-TF32 mode must be
within its 10-bit mantissa.  Also: apex_mlp_forward / apex_mlp_backward give the same result on the tensor-core and SIMT routes."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from apex_b200 import _capi
    return _capi.lib(), _capi


def _s():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("M", [128, 1000, 4096 + 37, 65536])
@pytest.mark.parametrize("passes", [3, 1])
def test_linear_forward_and_dx(M, passes):
    L, capi = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + passes)
    K = 256
    A = torch.randn(M, K, device="cuda", generator=g) * torch.exp(torch.randn(M, 1, device="cuda", generator=g))
    W = torch.randn(256, K, device="cuda", generator=g) / 16
    b = torch.randn(256, device="cuda", generator=g)
    mask = torch.randn(M, 256, device="cuda", generator=g)
    tol = 2e-6 if passes == 3 else 2e-3  # relative to sum |a||w|: a float32 FMA chain of length 256 sits at ~1e-6 too
    scale = (A.double().abs() @ W.double().abs().t()) + 1e-30
    # forward: relu(A W^T + b)
    C = torch.full((M, 256), float("nan"), device="cuda")
    capi.check(L.apex_tc3_linear(A.data_ptr(), K, M, K, W.data_ptr(), K, 1, b.data_ptr(), 1, None, 0, C.data_ptr(), 256, passes, _s()), "tc3_linear")
    ref = torch.relu(A.double() @ W.double().t() + b.double())
    err = ((C.double() - ref).abs() / scale).max().item()
    assert err < tol, err
    # dX: (A W) masked, W read through its transpose
    C2 = torch.full((M, 256), float("nan"), device="cuda")
    capi.check(L.apex_tc3_linear(A.data_ptr(), K, M, K, W.data_ptr(), 1, 256, None, 0, mask.data_ptr(), 256, C2.data_ptr(), 256, passes, _s()), "tc3_linear")
    ref2 = (A.double() @ W.double()) * (mask > 0)
    scale2 = (A.double().abs() @ W.double().abs()) + 1e-30
    err2 = ((C2.double() - ref2).abs() / scale2).max().item()
    assert err2 < tol, err2
    assert torch.equal(C2 == 0, ~(mask > 0) | (C2 == 0))


@pytest.mark.parametrize("R", [16, 5000 + 3, 65536])
@pytest.mark.parametrize("passes", [3, 1])
def test_outer_weight_gradient(R, passes):
    L, capi = _lib()
    g = torch.Generator(device="cuda").manual_seed(R + passes)
    A = torch.randn(R, 256, device="cuda", generator=g)
    B = torch.relu(torch.randn(R, 256, device="cuda", generator=g))
    C = torch.full((256, 256), 7.0, device="cuda")
    capi.check(L.apex_tc3_outer(A.data_ptr(), 256, B.data_ptr(), 256, 256, R, C.data_ptr(), 256, 0, passes, _s()), "tc3_outer")
    ref = A.double().t() @ B.double()
    scale = A.double().abs().t() @ B.double().abs() + 1e-30
    tol = 2e-6 if passes == 3 else 2e-3
    err = ((C.double() - ref).abs() / scale).max().item()
    assert err < tol, err
    capi.check(L.apex_tc3_outer(A.data_ptr(), 256, B.data_ptr(), 256, 256, R, C.data_ptr(), 256, 1, passes, _s()), "tc3_outer")
    err = ((C.double() - 2 * ref).abs() / scale).max().item()
    assert err < 2 * tol, err


def test_split_is_as_accurate_as_the_simt_kernel():
    """mode 3 against mode 0 on the same MLP forward + backward.  Both are float32 computations; the tensor-core route rounds
    more often (three products per k step, accumulator truncation in the MMA), measured ~2e-6 of max |y| against ~5e-7 for the
    FFMA chain — bounded here at 8x / 4e-6.  Every stage is compared with float64 applied to the kernel's OWN inputs of that
    stage (its h1 / h2 / dh2), so a ReLU mask that flips on a pre-activation within rounding of zero does not enter."""
    L, capi = _lib()
    rows, din, hid, dout = 8192, 50, 256, 10
    g = torch.Generator(device="cuda").manual_seed(5)
    f = dict(device="cuda", generator=g)
    x = torch.randn(rows, din, **f)
    w1, b1 = torch.randn(hid, din, **f) / 7, torch.randn(hid, **f) / 10
    w2, b2 = torch.randn(hid, hid, **f) / 16, torch.randn(hid, **f) / 10
    w3, b3 = torch.randn(dout, hid, **f) / 16, torch.randn(dout, **f) / 10
    dy = torch.randn(rows, dout, **f)
    w2d, b2d, w3d, b3d, dyd = (t.double() for t in (w2, b2, w3, b3, dy))
    errs = {}
    for mode in (0, 3):
        L.apex_set_tc_mode(mode)
        h1, h2, y = (torch.empty(rows, n, device="cuda") for n in (hid, hid, dout))
        capi.check(L.apex_mlp_forward(x.data_ptr(), rows, din, hid, dout, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                      w3.data_ptr(), b3.data_ptr(), h1.data_ptr(), h2.data_ptr(), y.data_ptr(), _s()), "fwd")
        dh2, dh1 = torch.empty(rows, hid, device="cuda"), torch.empty(rows, hid, device="cuda")
        gw1, gb1, gw2, gb2, gw3, gb3 = (torch.zeros_like(t) for t in (w1, b1, w2, b2, w3, b3))
        capi.check(L.apex_mlp_backward(x.data_ptr(), rows, din, hid, dout, w2.data_ptr(), w3.data_ptr(), h1.data_ptr(), h2.data_ptr(),
                                       dy.data_ptr(), dh2.data_ptr(), dh1.data_ptr(), gw1.data_ptr(), gb1.data_ptr(), gw2.data_ptr(),
                                       gb2.data_ptr(), gw3.data_ptr(), gb3.data_ptr(), _s()), "bwd")
        pre2 = h1.double() @ w2d.t() + b2d
        clear = pre2.abs() > 1e-4  # entries whose ReLU branch cannot depend on float32 rounding
        refs = {"h2": (torch.where(clear, h2.double(), torch.relu(pre2)), torch.relu(pre2)),
                "gw2": (gw2.double(), dh2.double().t() @ h1.double()),
                "dh1": (dh1.double(), (dh2.double() @ w2d) * (h1 > 0))}
        errs[mode] = {k: ((a - r).abs().max() / r.abs().max()).item() for k, (a, r) in refs.items()}
    L.apex_set_tc_mode(3)
    for k in errs[0]:
        assert errs[3][k] < max(8 * errs[0][k], 4e-6), (k, errs)


@pytest.mark.parametrize("dout", [10, 1])
@pytest.mark.parametrize("rows", [1024, 5000 + 3, 65536])
def test_head_kernels_match_float64(rows, dout):
    """csrc/mlp_head.cu: output layer forward and the fused backward (dh2, gW3, gb3 in one pass) against float64 and against
    the GEMM route they replace."""
    L, capi = _lib()
    din, hid = 50, 256
    g = torch.Generator(device="cuda").manual_seed(rows + dout)
    f = dict(device="cuda", generator=g)
    x = torch.randn(rows, din, **f)
    w1, b1 = torch.randn(hid, din, **f) / 7, torch.randn(hid, **f) / 10
    w2, b2 = torch.randn(hid, hid, **f) / 16, torch.randn(hid, **f) / 10
    w3, b3 = torch.randn(dout, hid, **f) / 16, torch.randn(dout, **f) / 10
    dy = torch.randn(rows, dout, **f)
    res = {}
    for head in (1, 0):
        L.apex_set_head_kernels(head)
        h1, h2, y = (torch.empty(rows, n, device="cuda") for n in (hid, hid, dout))
        capi.check(L.apex_mlp_forward(x.data_ptr(), rows, din, hid, dout, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                      w3.data_ptr(), b3.data_ptr(), h1.data_ptr(), h2.data_ptr(), y.data_ptr(), _s()), "fwd")
        dh2, dh1 = torch.empty(rows, hid, device="cuda"), torch.empty(rows, hid, device="cuda")
        gw1, gb1, gw2, gb2, gw3, gb3 = (torch.zeros_like(t) for t in (w1, b1, w2, b2, w3, b3))
        capi.check(L.apex_mlp_backward(x.data_ptr(), rows, din, hid, dout, w2.data_ptr(), w3.data_ptr(), h1.data_ptr(), h2.data_ptr(),
                                       dy.data_ptr(), dh2.data_ptr(), dh1.data_ptr(), gw1.data_ptr(), gb1.data_ptr(), gw2.data_ptr(),
                                       gb2.data_ptr(), gw3.data_ptr(), gb3.data_ptr(), _s()), "bwd")
        res[head] = (h2, y, dh2, gw3, gb3)
    L.apex_set_head_kernels(1)
    h2, y, dh2, gw3, gb3 = res[1]
    h2d, dyd, w3d = h2.double(), dy.double(), w3.double()
    refs = {"y": (y, h2d @ w3d.t() + b3.double()), "dh2": (dh2, (dyd @ w3d) * (h2 > 0)), "gw3": (gw3, dyd.t() @ h2d), "gb3": (gb3, dyd.sum(0))}
    for k, (a, r) in refs.items():
        err = ((a.double() - r).abs().max() / r.abs().max()).item()
        assert err < (3e-5 if k == "gb3" else 3e-6), (k, err)  # gb3: a float32 sum of `rows` terms against its (cancelling) total
    assert torch.equal(res[0][0], h2)
    for a, b in zip(res[0][1:], res[1][1:]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))


@pytest.mark.parametrize("passes", [3, 1])
def test_first_layer_shapes(passes):
    """The 50-wide observation layer: forward with K = 50 (rows of 200 bytes: element-wise loads, k padded to 64 on chip) and the
    weight gradient dW1 [256, 50] = dh1^T x through the narrow (N = 64) variant of the split-k kernel."""
    L, capi = _lib()
    g = torch.Generator(device="cuda").manual_seed(11 + passes)
    M, K = 7000 + 5, 50
    x = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(256, K, device="cuda", generator=g) / 7
    b = torch.randn(256, device="cuda", generator=g)
    tol = 2e-6 if passes == 3 else 2e-3
    C = torch.full((M, 256), float("nan"), device="cuda")
    capi.check(L.apex_tc3_linear(x.data_ptr(), K, M, K, W.data_ptr(), K, 1, b.data_ptr(), 1, None, 0, C.data_ptr(), 256, passes, _s()), "tc3_linear")
    ref = torch.relu(x.double() @ W.double().t() + b.double())
    scale = x.double().abs() @ W.double().abs().t() + b.double().abs() + 1e-30
    err = ((C.double() - ref).abs() / scale).max().item()
    assert err < tol, err
    dh = torch.randn(M, 256, device="cuda", generator=g)
    G = torch.full((256, K), 3.0, device="cuda")
    capi.check(L.apex_tc3_outer(dh.data_ptr(), 256, x.data_ptr(), K, K, M, G.data_ptr(), K, 0, passes, _s()), "tc3_outer")
    refg = dh.double().t() @ x.double()
    scaleg = dh.double().abs().t() @ x.double().abs() + 1e-30
    errg = ((G.double() - refg).abs() / scaleg).max().item()
    assert errg < tol, errg


def test_parameters_at_any_float_offset():
    """The critic's parameters start 8 bytes into a 16-byte line of the flat parameter buffer (81,418 actor floats precede them):
    weights, biases and weight gradients at odd float offsets must take the tensor-core / streaming routes and give the same
    result as 16-byte aligned copies."""
    L, capi = _lib()
    rows, din, hid, dout = 4096, 50, 256, 1
    g = torch.Generator(device="cuda").manual_seed(21)
    f = dict(device="cuda", generator=g)
    x, dy = torch.randn(rows, din, **f), torch.randn(rows, dout, **f)
    shapes = [(hid, din), (hid,), (hid, hid), (hid,), (dout, hid), (dout,)]
    vals = [torch.randn(*s, **f) / 8 for s in shapes]
    out = {}
    for off in (0, 2):
        n = sum(v.numel() for v in vals)
        flat, grad = torch.zeros(n + 8, device="cuda"), torch.zeros(n + 8, device="cuda")
        ptr, views, gviews = off, [], []
        for v in vals:
            views.append(flat[ptr:ptr + v.numel()].view_as(v)); views[-1].copy_(v)
            gviews.append(grad[ptr:ptr + v.numel()].view_as(v))
            ptr += v.numel()
        assert (views[0].data_ptr() % 16 == 0) == (off == 0)
        w1, b1, w2, b2, w3, b3 = views
        h1, h2, y = (torch.empty(rows, k, device="cuda") for k in (hid, hid, dout))
        capi.check(L.apex_mlp_forward(x.data_ptr(), rows, din, hid, dout, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                      w3.data_ptr(), b3.data_ptr(), h1.data_ptr(), h2.data_ptr(), y.data_ptr(), _s()), "fwd")
        dh2, dh1 = torch.empty(rows, hid, device="cuda"), torch.empty(rows, hid, device="cuda")
        gw1, gb1, gw2, gb2, gw3, gb3 = gviews
        capi.check(L.apex_mlp_backward(x.data_ptr(), rows, din, hid, dout, w2.data_ptr(), w3.data_ptr(), h1.data_ptr(), h2.data_ptr(),
                                       dy.data_ptr(), dh2.data_ptr(), dh1.data_ptr(), gw1.data_ptr(), gb1.data_ptr(), gw2.data_ptr(),
                                       gb2.data_ptr(), gw3.data_ptr(), gb3.data_ptr(), _s()), "bwd")
        out[off] = [t.clone() for t in (y, dh1, gw1, gb1, gw2, gb2, gw3, gb3)]
    for a, b in zip(out[0], out[2]):
        assert torch.allclose(a, b, rtol=1e-5, atol=2e-6 * float(a.abs().max())), float((a - b).abs().max())
