"""CPU checks of host-side arithmetic the GPU path relies on (no kernel is called here).

  * the factorisation behind the SE blocks' one-pass BatchNorm-backward reduce (csrc/bwd_kernels.cu:
    se_bn_bwd_reduce_kernel / se_bn_bwd_combine_kernel) and the constant folding of bn_bwd_apply2_kernel, in fp64 torch
    against autograd of the reference's own expression (block_types.py:72-83 in front of BatchNorm2d + Hardswish);
  * bench.py's algorithmic-byte accounting for the entry points whose GB/s the kernel table prints."""
import importlib.util
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_se_bn_backward_factorisation_matches_autograd():
    torch.manual_seed(0)
    B, C, P = 3, 8, 50
    z = torch.randn(B, C, P, dtype=torch.double, requires_grad=True)            # raw depthwise output, [B, C, pixels]
    gamma, beta = torch.rand(C, dtype=torch.double) + 0.5, torch.randn(C, dtype=torch.double)
    w1, w2 = torch.randn(4, C, dtype=torch.double) * 0.3, torch.randn(C, 4, dtype=torch.double) * 0.3
    eps = 1e-3
    mean, var = z.mean((0, 2)), z.var((0, 2), unbiased=False)
    invstd = (var + eps).rsqrt()
    xhat = (z - mean[None, :, None]) * invstd[None, :, None]
    v = xhat * gamma[None, :, None] + beta[None, :, None]
    a = torch.nn.functional.hardswish(v)
    gate = torch.sigmoid(torch.relu(a.mean(2) @ w1.T) @ w2.T)                   # SqueezeExcitation: mean -> fc1 -> ReLU -> fc2 -> Sigmoid
    out = a * gate[:, :, None]                                                   # scale * input
    dp = torch.randn_like(out)
    (dz_ref,) = torch.autograd.grad(out, z, dp)

    with torch.no_grad():
        scale, shift = gamma * invstd, beta - mean * gamma * invstd
        vv = z * scale[None, :, None] + shift[None, :, None]
        f = torch.nn.functional.hardswish(vv)
        d = torch.where(vv < -3, torch.zeros_like(vv), torch.where(vv <= 3, (2 * vv + 3) / 6, torch.ones_like(vv)))
        zc = z - mean[None, :, None]
        # pass 1 (one walk over dp, z): dgate and the four pixel sums per (sample, channel)
        dgate = (dp * f).sum(2)
        A1, A2, E1, E2 = (dp * d).sum(2), (dp * d * zc).sum(2), d.sum(2), (d * zc).sum(2)
        # SE MLP backward -> dpool (gradient w.r.t. the pooled mean, already divided by the pixel count)
        hidden = torch.relu(a.mean(2) @ w1.T)
        du2 = dgate * gate * (1 - gate)
        du1 = (du2 @ w2) * (hidden > 0)
        dpool = (du1 @ w1) / P
        # combine over the batch: what the two-tensor reduce pass would have produced
        s1 = (gate * A1 + dpool * E1).sum(0)
        s2 = (gate * A2 + dpool * E2).sum(0) * invstd
        direct = (dp * gate[:, :, None] + dpool[:, :, None]) * d
        assert torch.allclose(s1, direct.sum((0, 2)), rtol=1e-12, atol=1e-12)
        assert torch.allclose(s2, (direct * zc).sum((0, 2)) * invstd, rtol=1e-12, atol=1e-12)
        # apply pass with the folded constants of bn_bwd_apply2_kernel
        M = B * P
        c1, c2 = s1 / M, s2 / M
        alpha = -scale * c2 * invstd
        beta2 = -scale * c1 - alpha * mean
        dz = scale[None, :, None] * direct + alpha[None, :, None] * z + beta2[None, :, None]
    assert torch.allclose(dz, dz_ref, rtol=1e-9, atol=1e-11), (dz - dz_ref).abs().max()


def _bench():
    spec = importlib.util.spec_from_file_location("_eat_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_bench_algorithmic_bytes_of_the_reduce_entry_points():
    b = _bench()
    B, P, C = 256, 128, 960
    n = B * P * C * 4
    assert b.algo_bytes("eat_se_bn_bwd_reduce", (1, 2, 3, 4, 5, 2, 7, 8, 3, 0, B, P, C, 0)) == 2 * n + 3 * 4 * B * C * 4
    assert b.algo_bytes("eat_se_bn_bwd_combine", (1, 3, 2, 3, 4, B, C, 7, 8, 0)) == (3 * 4 + 2) * B * C * 4
    assert b.algo_bytes("eat_se_bwd_reduce", (1, 2, 3, 4, 2, 5, 0, B, P, C, 0)) == 2 * n
    assert b.algo_bytes("eat_bn_act_pool", (1, 2, 3, 2, 4, 0.5, 1, B, P, C, 0)) == n // 2           # bf16 storage
    assert b.algo_bytes("eat_bn_bwd_reduce", (1, 0, 0, 2, 3, 4, 5, 6, 2, 0, B, P, C, 7, 8, 0)) == 2 * n
    assert b.algo_bytes("eat_bn_bwd_apply", (0, 0, 1, 2, 3, 4, 5, 6, 2, 7, 8, 9, 0, B, P, C, 0)) == 2 * n   # no gA tensor
    assert b.algo_bytes("eat_dw_conv_dgrad_bnred", (1, 2, 0, 3, 4, 5, 6, 7, 8, 2, 9, 10, 0, B, 64, 500, 64, 3, 2, 0)) == \
        B * 64 * 4 * (64 * 500 * 2 + 32 * 250)
    assert b.algo_bytes("eat_not_an_entry_point", ()) is None
