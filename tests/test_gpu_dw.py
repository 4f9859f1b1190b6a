"""GPU parity of the depthwise-convolution kernels through the C ABI (csrc/dw_slide.cu, conv_kernels.cu,
bwd_kernels.cu) against a float64 torch reference (F.conv2d with groups = C and its autograd), over the layer
shapes of mn10 plus ragged cases: widths that are not a multiple of the strip, channel counts that do not fill a
CTA, single-row / single-column images, batch 1.  fp32 storage: 1e-5 of the output scale (exact fp32 FMAs, only
the summation order differs); bf16 storage: 2e-2 (bf16 rounding of inputs and outputs)."""
import pytest
import torch
import torch.nn.functional as Fn

from efficientat_b200._lib import lib

pytestmark = pytest.mark.gpu

CASES = [  # (B, F, T, C, k, stride)
    (2, 64, 100, 16, 3, 1), (2, 64, 101, 64, 3, 2), (3, 32, 50, 72, 3, 1), (2, 32, 51, 72, 5, 2), (2, 16, 25, 120, 5, 1),
    (2, 16, 26, 240, 3, 2), (2, 8, 13, 200, 3, 1), (2, 8, 13, 672, 5, 2), (2, 4, 7, 960, 5, 1), (1, 1, 9, 32, 3, 1),
    (1, 9, 1, 32, 5, 1), (1, 2, 2, 40, 3, 2), (1, 5, 3, 2560, 5, 2), (2, 37, 41, 24, 3, 1), (1, 70, 33, 8, 5, 1),
]


def _act(x, code):
    return torch.relu(x) if code == 1 else (Fn.hardswish(x) if code == 2 else x)


def _st():
    return torch.cuda.current_stream().cuda_stream


def _setup(B, F, T, C, k, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(B, F, T, C, device="cuda", generator=g).to(dtype)
    w = torch.randn(C, 1, k, k, device="cuda", generator=g) * 0.3
    wt = torch.empty(k * k, C, device="cuda")
    lib().dw_repack(w.data_ptr(), wt.data_ptr(), C, k, _st())
    return g, x, w, wt


def _conv_ref(a_nhwc, w, k, s):
    return Fn.conv2d(a_nhwc.permute(0, 3, 1, 2), w.double(), None, s, (k - 1) // 2, 1, w.shape[0]).permute(0, 2, 3, 1)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("in_act", [-1, 0, 1, 2])
def test_dw_forward_training_mode(dtype, in_act):
    """BN affine + activation of the producing layer applied on load, raw output, fp64 batch statistics."""
    L = lib()
    code = 0 if dtype == torch.float32 else 1
    tol = 1e-5 if code == 0 else 2e-2
    for idx, (B, F, T, C, k, s) in enumerate(CASES):
        g, x, w, wt = _setup(B, F, T, C, k, dtype, idx)
        sc = torch.rand(2, C, device="cuda", generator=g) + 0.5
        sc[1] -= 1.0
        a = x.double()
        if in_act >= 0:
            a = _act(a * sc[0].double() + sc[1].double(), in_act)
        ref = _conv_ref(a, w, k, s)
        Fo, To = ref.shape[1], ref.shape[2]
        out = torch.full((B, Fo, To, C), float("nan"), device="cuda", dtype=dtype)
        stats = torch.zeros(2, C, device="cuda", dtype=torch.float64)
        L.dw_conv_fwd(x.data_ptr(), wt.data_ptr(), out.data_ptr(), code, B, F, T, C, k, s,
                      sc[0].data_ptr() if in_act >= 0 else 0, sc[1].data_ptr() if in_act >= 0 else 0, max(in_act, 0),
                      0, 0, 0, 0, stats[0].data_ptr(), stats[1].data_ptr(), _st())
        scale = ref.abs().max().item() + 1e-6
        err = (out.double() - ref).abs().max().item()
        assert err <= tol * scale, (idx, (B, F, T, C, k, s), err, scale)
        o = out.double()       # statistics are taken over the stored (rounded) values' fp32 accumulators
        n = B * Fo * To
        assert torch.allclose(stats[0], ref.sum((0, 1, 2)), rtol=0, atol=(1e-4 if code == 0 else 2e-2) * scale * n), idx
        assert torch.allclose(stats[1], (ref * ref).sum((0, 1, 2)), rtol=0, atol=(1e-4 if code == 0 else 3e-2) * scale * scale * n), idx
        del o


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_dw_forward_eval_mode(dtype, act):
    """folded BN + activation epilogue and the squeeze-excitation pooling sums."""
    L = lib()
    code = 0 if dtype == torch.float32 else 1
    tol = 1e-5 if code == 0 else 2e-2
    for idx, (B, F, T, C, k, s) in enumerate(CASES):
        g, x, w, wt = _setup(B, F, T, C, k, dtype, 100 + idx)
        sc = torch.rand(2, C, device="cuda", generator=g) + 0.5
        sc[1] -= 1.0
        ref = _act(_conv_ref(x.double(), w, k, s) * sc[0].double() + sc[1].double(), act)
        Fo, To = ref.shape[1], ref.shape[2]
        out = torch.full((B, Fo, To, C), float("nan"), device="cuda", dtype=dtype)
        pool = torch.zeros(B, C, device="cuda")
        L.dw_conv_fwd(x.data_ptr(), wt.data_ptr(), out.data_ptr(), code, B, F, T, C, k, s, 0, 0, 0,
                      sc[0].data_ptr(), sc[1].data_ptr(), act, pool.data_ptr(), 0, 0, _st())
        scale = ref.abs().max().item() + 1e-6
        err = (out.double() - ref).abs().max().item()
        assert err <= tol * scale, (idx, (B, F, T, C, k, s), err, scale)
        perr = (pool.double() - ref.sum((1, 2))).abs().max().item()
        assert perr <= (1e-4 if code == 0 else 2e-2) * scale * Fo * To, (idx, perr)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dw_backward_data_and_weight_gradients(dtype):
    """dgrad (+ residual-gradient add for stride 1) and wgrad (input transform on load) against autograd."""
    L = lib()
    code = 0 if dtype == torch.float32 else 1
    tol = 2e-5 if code == 0 else 2e-2
    for idx, (B, F, T, C, k, s) in enumerate(CASES):
        g, x, w, wt = _setup(B, F, T, C, k, dtype, 200 + idx)
        sc = torch.rand(2, C, device="cuda", generator=g) + 0.5
        sc[1] -= 1.0
        a = _act(x.double() * sc[0].double() + sc[1].double(), 2).requires_grad_(True)
        wd = w.double().requires_grad_(True)
        z = Fn.conv2d(a.permute(0, 3, 1, 2), wd, None, s, (k - 1) // 2, 1, C).permute(0, 2, 3, 1)
        Fo, To = z.shape[1], z.shape[2]
        dz = torch.randn(B, Fo, To, C, device="cuda", generator=g).to(dtype)
        ga, gw = torch.autograd.grad(z, (a, wd), dz.double())
        res = torch.randn(B, F, T, C, device="cuda", generator=g).to(dtype) if s == 1 else None
        din = torch.full((B, F, T, C), float("nan"), device="cuda", dtype=dtype)
        L.dw_conv_dgrad(dz.data_ptr(), wt.data_ptr(), 0, res.data_ptr() if res is not None else 0, din.data_ptr(), code,
                        B, F, T, C, k, s, _st())
        want = ga + (res.double() if res is not None else 0)
        scale = want.abs().max().item() + 1e-6
        err = (din.double() - want).abs().max().item()
        assert err <= tol * scale, ("dgrad", idx, (B, F, T, C, k, s), err, scale)
        dw = torch.zeros(C, 1, k, k, device="cuda")
        L.dw_conv_wgrad(dz.data_ptr(), x.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(), 2, dw.data_ptr(), 0, code,
                        B, F, T, C, k, s, _st())
        wscale = gw.abs().max().item() + 1e-6
        werr = (dw.double() - gw).abs().max().item()
        assert werr <= (1e-4 if code == 0 else 2e-2) * wscale, ("wgrad", idx, (B, F, T, C, k, s), werr, wscale)

