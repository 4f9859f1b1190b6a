"""(Named zz_ so that it is collected after the model-level files: it is the newest file of the suite, and the files before
it then run in the process state they were validated in.)

Kernel-level checks of the BatchNorm-backward reduce passes against fp64 torch: the plain two-tensor reduce
(eat_bn_bwd_reduce with gate / dpool composition) and the SE-block variant that takes the squeeze-excitation sum and the four
BatchNorm pixel sums in ONE pass (eat_se_bn_bwd_reduce) and combines them over the batch once dpool is known
(eat_se_bn_bwd_combine).  Autograd of block_types.py:72-83 (`scale * input`) in front of a training-mode BatchNorm +
activation (block_types.py:150-162)."""
import pytest
import torch

from efficientat_b200._lib import lib

pytestmark = pytest.mark.gpu
ACT = {"none": 0, "relu": 1, "hswish": 2}


def _act(v, act):
    if act == "relu":
        return torch.relu(v), (v > 0).double()
    if act == "hswish":
        f = v * torch.clamp(v + 3, 0, 6) / 6
        d = torch.where(v < -3, torch.zeros_like(v), torch.where(v <= 3, (2 * v + 3) / 6, torch.ones_like(v)))
        return f, d
    return v, torch.ones_like(v)


def _case(B, P, C, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(B, P, C, generator=g) * 1.5
    dp = torch.randn(B, P, C, generator=g)
    scale = torch.rand(C, generator=g) + 0.5
    shift = torch.randn(C, generator=g) * 0.3
    mean = torch.randn(C, generator=g) * 0.2
    invstd = torch.rand(C, generator=g) + 0.5
    gate = torch.rand(B, C, generator=g)
    dpool = torch.randn(B, C, generator=g) * 0.1
    z, dp = z.to(dtype), dp.to(dtype)
    return [t.cuda().contiguous() for t in (z, dp, scale, shift, mean, invstd, gate, dpool)]


def _reference(z, dp, scale, shift, mean, invstd, gate, dpool, act):
    z, dp = z.double(), dp.double()
    v = z * scale.double() + shift.double()
    f, d = _act(v, act)
    dgate = (dp * f).sum(1)
    dy = (dp * gate.double()[:, None, :] + dpool.double()[:, None, :]) * d
    s1 = dy.sum((0, 1))
    s2 = (dy * (z - mean.double())).sum((0, 1)) * invstd.double()
    return dgate, s1, s2


def _close(got, want, rel):
    scale = want.abs().max().item() + 1e-12
    err = (got.double() - want).abs().max().item()
    assert err <= rel * scale, (err, scale)


# (B, P, C): several pixel slots per channel vector (shared-memory flush) / one owner per vector / more vectors than threads
SHAPES = [(3, 37, 72), (5, 2000, 120), (2, 130, 960), (2, 50, 1536), (1, 1, 8), (4, 504, 672)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", ["relu", "hswish"])
@pytest.mark.parametrize("shape", SHAPES)
def test_se_bn_bwd_fused_reduce_matches_fp64_and_the_two_pass_kernels(shape, act, dtype):
    B, P, C = shape
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    z, dp, scale, shift, mean, invstd, gate, dpool = _case(B, P, C, dtype)
    code = 1 if dtype == torch.bfloat16 else 0
    want_dgate, want_s1, want_s2 = _reference(z, dp, scale, shift, mean, invstd, gate, dpool, act)
    for parts in (1, 3, 7):
        dgate = torch.zeros(B, C, device="cuda")
        part = torch.full((parts, 4, B, C), 7.0e3, device="cuda")              # every slice must be written: a stale 7e3 wrecks the sums
        L.se_bn_bwd_reduce(dp.data_ptr(), z.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(), ACT[act],
                           dgate.data_ptr(), part.data_ptr(), parts, code, B, P, C, st)
        s = torch.zeros(2, C, device="cuda", dtype=torch.float64)
        L.se_bn_bwd_combine(part.data_ptr(), parts, gate.data_ptr(), dpool.data_ptr(), invstd.data_ptr(), B, C,
                            s[0].data_ptr(), s[1].data_ptr(), st)
        torch.cuda.synchronize()
        assert not (part == 7.0e3).any()
        _close(dgate.cpu(), want_dgate.cpu(), 5e-5)
        _close(s[0].cpu(), want_s1.cpu(), 5e-5)
        _close(s[1].cpu(), want_s2.cpu(), 5e-5)
    # the kernels it replaces, on the same data
    dgate2 = torch.zeros(B, C, device="cuda")
    L.se_bwd_reduce(dp.data_ptr(), z.data_ptr(), scale.data_ptr(), shift.data_ptr(), ACT[act], dgate2.data_ptr(), code, B, P, C, st)
    s2 = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    L.bn_bwd_reduce(dp.data_ptr(), gate.data_ptr(), dpool.data_ptr(), z.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                    mean.data_ptr(), invstd.data_ptr(), ACT[act], code, B, P, C, s2[0].data_ptr(), s2[1].data_ptr(), st)
    torch.cuda.synchronize()
    _close(dgate2.cpu(), want_dgate.cpu(), 5e-5)
    _close(s2[0].cpu(), want_s1.cpu(), 5e-5)
    _close(s2[1].cpu(), want_s2.cpu(), 5e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", ["none", "relu", "hswish"])
@pytest.mark.parametrize("shape", SHAPES)
def test_bn_act_pool_matches_fp64(shape, act, dtype):
    """pool[b,c] += mul * sum_p act(z*scale+shift): the SE squeeze of a training step (block_types.py:73) and the global
    average pool of the head (mn/model.py:216).  Runs whichever kernel generation EAT_POOL selects (default: see
    kPoolV2Default in csrc/conv_kernels.cu); scripts/gpu_runs run the file under both."""
    B, P, C = shape
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    z, _, scale, shift, *_ = _case(B, P, C, dtype, seed=3)
    code = 1 if dtype == torch.bfloat16 else 0
    pool = torch.full((B, C), 0.25, device="cuda")
    L.bn_act_pool(z.data_ptr(), scale.data_ptr(), shift.data_ptr(), ACT[act], pool.data_ptr(), 0.5, code, B, P, C, st)
    torch.cuda.synchronize()
    f, _ = _act(z.double() * scale.double() + shift.double(), act)
    want = 0.25 + 0.5 * f.sum(1)
    _close(pool.cpu(), want.cpu(), 5e-5)


@pytest.mark.parametrize("act", ["relu", "hswish"])
@pytest.mark.parametrize("shape", [(3, 9, 21, 72, 5), (2, 16, 50, 64, 3), (2, 8, 13, 672, 5), (1, 2, 3, 8, 3), (2, 64, 100, 24, 3)])
def test_dw_dgrad_stride2_with_bn_reduce_epilogue(shape, act):
    """eat_dw_conv_dgrad_bnred: the same din as eat_dw_conv_dgrad, bit for bit, and the sums eat_bn_bwd_reduce(gA = din)
    yields (also checked against fp64 torch) -- the expand-stage BatchNorm of a stride-2 InvertedResidual."""
    B, F, T, C, k = shape
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator().manual_seed(11)
    pad = (k - 1) // 2
    Fo, To = (F + 2 * pad - k) // 2 + 1, (T + 2 * pad - k) // 2 + 1
    dz = torch.randn(B, Fo, To, C, generator=g).cuda()
    w = (torch.randn(C, 1, k, k, generator=g) * 0.3).cuda()
    z = (torch.randn(B, F, T, C, generator=g) * 1.5).cuda()
    scale, shift = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.3).cuda()
    mean, invstd = (torch.randn(C, generator=g) * 0.2).cuda(), (torch.rand(C, generator=g) + 0.5).cuda()
    wt = torch.empty(k * k, C, device="cuda")
    L.dw_repack(w.data_ptr(), wt.data_ptr(), C, k, st)
    din0 = torch.empty(B, F, T, C, device="cuda")
    L.dw_conv_dgrad(dz.data_ptr(), wt.data_ptr(), 0, 0, din0.data_ptr(), 0, B, F, T, C, k, 2, st)
    s0 = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    L.bn_bwd_reduce(din0.data_ptr(), 0, 0, z.data_ptr(), scale.data_ptr(), shift.data_ptr(), mean.data_ptr(),
                    invstd.data_ptr(), ACT[act], 0, B, F * T, C, s0[0].data_ptr(), s0[1].data_ptr(), st)
    din1 = torch.full((B, F, T, C), 7.0e3, device="cuda")
    s1 = torch.zeros(2, C, device="cuda", dtype=torch.float64)
    L.dw_conv_dgrad_bnred(dz.data_ptr(), wt.data_ptr(), 0, din1.data_ptr(), z.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                          mean.data_ptr(), invstd.data_ptr(), ACT[act], s1[0].data_ptr(), s1[1].data_ptr(), 0, B, F, T, C, k, 2, st)
    torch.cuda.synchronize()
    assert torch.equal(din0, din1)
    # fp64 reference of the data gradient (conv_transpose of the depthwise kernel) and of the reduce
    ref = torch.nn.functional.conv_transpose2d(dz.double().permute(0, 3, 1, 2), w.double(), stride=2, padding=pad, groups=C,
                                               output_padding=(F - ((Fo - 1) * 2 - 2 * pad + k), T - ((To - 1) * 2 - 2 * pad + k)))
    ref = ref.permute(0, 2, 3, 1)
    assert (din1.double() - ref).abs().max().item() <= 1e-5 * (ref.abs().max().item() + 1e-12)
    v = z.double() * scale.double() + shift.double()
    _, d = _act(v, act)
    gd = ref * d
    want1 = gd.sum((0, 1, 2))
    want2 = (gd * (z.double() - mean.double())).sum((0, 1, 2)) * invstd.double()
    for got in (s0, s1):
        _close(got[0].cpu(), want1.cpu(), 5e-5)
        _close(got[1].cpu(), want2.cpu(), 5e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("compose", ["plain", "gate_dpool", "dpool_only"])
@pytest.mark.parametrize("act", ["none", "relu", "hswish"])
@pytest.mark.parametrize("shape", [(3, 37, 72), (2, 130, 960), (2, 50, 1536), (1, 1, 8), (5, 2001, 16)])
def test_bn_bwd_apply_matches_fp64(shape, act, compose, dtype):
    """dz = scale * (dy - c1 - xhat * c2), dy = (gA * gate + dpool) * act'(z*scale+shift): pass 2 of the BatchNorm backward
    (autograd of nn.BatchNorm2d in training mode followed by Hardswish / ReLU).  Runs whichever kernel generation
    EAT_BN_APPLY selects (default: kBnApplyV2Default in csrc/bwd_kernels.cu)."""
    B, P, C = shape
    L = lib()
    st = torch.cuda.current_stream().cuda_stream
    z, gA, scale, shift, mean, invstd, gate, dpool = _case(B, P, C, dtype, seed=5)
    g = torch.Generator().manual_seed(9)
    c1, c2 = (torch.randn(C, generator=g) * 0.1).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
    code = 1 if dtype == torch.bfloat16 else 0
    p = lambda t: t.data_ptr()
    use_gA = compose != "dpool_only"
    use_gate = compose == "gate_dpool"
    use_dp = compose != "plain"
    dz = torch.full((B, P, C), 7.0e3, device="cuda", dtype=dtype)
    L.bn_bwd_apply(p(gA) if use_gA else 0, p(gate) if use_gate else 0, p(dpool) if use_dp else 0, p(z), p(scale), p(shift),
                   p(mean), p(invstd), ACT[act], p(c1), p(c2), p(dz), code, B, P, C, st)
    torch.cuda.synchronize()
    zd = z.double()
    _, d = _act(zd * scale.double() + shift.double(), act)
    gg = gA.double() if use_gA else torch.zeros_like(zd)
    if use_gate:
        gg = gg * gate.double()[:, None, :]
    if use_dp:
        gg = gg + dpool.double()[:, None, :]
    dy = gg * d
    xhat = (zd - mean.double()) * invstd.double()
    want = scale.double() * (dy - c1.double() - xhat * c2.double())
    tol = 1e-5 if dtype == torch.float32 else 1e-2        # bf16: the stored result is rounded to 8 bits of mantissa
    assert torch.isfinite(dz.float()).all()
    _close(dz.cpu(), want.cpu(), tol)
