"""GPU parity of the DyMN eval forward (ContextGen, DynamicConv, DyReLU-B, CoordAtt fused kernels) against the
reference's golden vectors: logits within 1e-3 max-abs (fp32 activation storage), top-10 identical up to
near-ties, per-block feature maps against the oracle."""
import numpy as np
import pytest
import torch

from oracle import net_oracle
from tests.util import report, NETS, build_model, fmap_digest, golden, net_inputs, topk_match

pytestmark = pytest.mark.gpu


def _layer_report(tag, model, spec):
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        _, _, ref = net_oracle.dymn_forward(sd, spec, width_mult=NETS[tag][1], temperature=30.0, return_fmaps=True)
        _, got = model(spec.cuda(), return_fmaps=True)
    return "\n".join(f"fmap {i:2d} {tuple(r.shape)} max-abs err {(g.float().cpu() - r).abs().max().item():.3e} "
                     f"(ref abs-max {r.abs().max().item():.3e})" for i, (r, g) in enumerate(zip(ref, got)))


@pytest.mark.parametrize("gemm", ["simt", "auto"])
@pytest.mark.parametrize("tag", ["dymn10", "dymn04", "dymn20", "dymn20_10s"])
def test_dymn_eval_fp32_matches_reference_vectors(tag, gemm):
    g = golden(tag)
    model = build_model(tag).cuda().eval()
    model.engine().gemm_impl = gemm
    spec, _ = net_inputs(tag)
    with torch.no_grad():
        logits, feat = model(spec.cuda())
    logits, feat = logits.cpu().numpy(), feat.cpu().numpy()
    err = np.abs(logits - g["eval_logits"]).max()
    report(f"[parity] {tag} gemm={gemm}: logit max-abs err {err:.3e}")
    if not err < 1e-3:
        pytest.fail(f"logit max-abs err {err}\n" + _layer_report(tag, model, spec))
    assert np.abs(feat - g["eval_feat"]).max() < 2e-3
    assert topk_match(logits, g["eval_logits"], 10, tie_tol=2e-4)


def test_dymn_eval_fmaps_and_temperature():
    tag = "dymn10"
    g = golden(tag)
    model = build_model(tag).cuda().eval()
    spec, _ = net_inputs(tag)
    with torch.no_grad():
        _, fmaps = model(spec.cuda(), return_fmaps=True)
    d = fmap_digest([f.float().cpu().contiguous() for f in fmaps])
    assert d.shape == g["eval_fmaps"].shape
    assert np.abs(d - g["eval_fmaps"]).max() < 5e-3, _layer_report(tag, model, spec)
    # update_params(epoch) changes the attention temperature (dy_block.py:133-139) and therefore the output
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model.update_params(12)          # T = 30 - 12 = 18 (a much lower T leaves the BN calibration far behind)
    t = model.layers[0].depth_conv.temperature
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        logits, _ = model(spec.cuda())
        ref, _ = net_oracle.dymn_forward(sd, spec, temperature=t)
    assert t == 18.0
    assert (logits.cpu() - ref).abs().max() < 1e-3 * max(1.0, ref.abs().max().item())
    assert (logits.cpu() - torch.from_numpy(g["eval_logits"])).abs().max() > 1e-3     # the temperature matters


def test_dymn_replace_se_variant_and_bf16():
    from efficientat_b200.models.dymn.model import get_model
    from efficientat_b200.synth import synth_state_
    torch.manual_seed(0)
    m = synth_state_(get_model(width_mult=0.4, use_dy_blocks="replace_se", verbose=False), seed=3).cuda().eval()
    spec = net_inputs("dymn04")[0]
    with torch.no_grad():
        logits, feat = m(spec.cuda())
    assert logits.shape == (2, 527) and torch.isfinite(logits).all()
    mb = build_model("dymn10", precision="bf16").cuda().eval()
    with torch.no_grad():
        lb, _ = mb(net_inputs("dymn10")[0].cuda())
    assert np.abs(lb.cpu().numpy() - golden("dymn10")["eval_logits"]).max() < 8e-2


@pytest.mark.parametrize("gemm", ["simt", "auto"])
@pytest.mark.parametrize("tag", ["dymn04", "dymn10", "dymn20"])
def test_dymn_train_step_matches_reference_vectors(tag, gemm):
    """batch-statistics forward + hand-written backward of the dynamic blocks vs the reference's autograd:
    loss, logits, every parameter's gradient norm and samples, BatchNorm running statistics.

    gemm = 'simt': every GEMM -- including the DynamicConv 1x1 forward, data gradient and per-sample weight gradient,
    which then follow the reference's own order (materialised per-sample kernels, dy_block.py:111-127) -- runs in exact
    fp32 on CUDA cores.  This is the independent implementation that shows the wide band of the tensor-core mode is
    arithmetic noise and not a bug: EVERY tensor must match to 5e-3 in norm (2e-2 for the attention-logit layers
    `*.residuals.0.*`, whose gradients are differences of nearly equal inner products <S_b, W_k> divided by the
    temperature 30 and ~100x smaller than every other gradient).
    gemm = 'auto': tcgen05 path (fp32 storage, bf16x3 products, ~2^-16 per product): tolerances as in
    tests/test_gpu_mn_train.py; the attention-logit gradients inherit the product noise of S_b amplified by that
    cancellation, hence their 20-30 % band.  Hardswish' jumps by 0.5 at +-3: with B = 2 a context-generator BatchNorm
    channel sees ~40 elements, so ONE pre-activation within product noise of a kink moves that channel's gradient by
    a few percent; up to two of the ~365 tensors may therefore leave the tight band in this mode, none the wide one."""
    g = golden(tag)
    model = build_model(tag).cuda().train()
    model.engine().gemm_impl = gemm
    model.classifier[4].p = 0.0
    model.engine().dropout_p = 0.0
    spec, y = net_inputs(tag)
    logits, _ = model(spec.cuda())
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y.cuda())
    loss.backward()
    lerr = np.abs(logits.detach().cpu().numpy() - g["train_logits"]).max()
    assert lerr < 1e-3
    assert abs(loss.item() - float(g["train_loss"])) < 2e-5
    params = dict(model.named_parameters())
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == set(params)
    exact = gemm == "simt"
    bad, very_bad = [], []
    worst = {"res": 0.0, "other": 0.0}
    for i, n in enumerate(names):
        gr = params[n].grad
        assert gr is not None, n
        gr = gr.detach().float().cpu()
        gn, ref = gr.double().norm().item(), g["grad_norm"][i]
        idx = torch.linspace(0, gr.numel() - 1, 4).long()
        samp = gr.flatten()[idx].numpy()
        is_res = ".residuals." in n
        if exact:
            ntol, stol = (2e-2, 5e-2) if is_res else (5e-3, 2e-2)
        else:
            ntol, stol = (0.2, 0.3) if is_res else (3e-2, 8e-2)
        atol = 5e-7 if is_res else 1e-8          # attention-logit gradients are O(1e-6): absolute floor
        if ref > 100 * atol:
            worst["res" if is_res else "other"] = max(worst["res" if is_res else "other"], abs(gn - ref) / ref)
        ok = abs(gn - ref) <= ntol * ref + 10 * atol and \
            np.abs(samp - g["grad_samples"][i]).max() <= stol * max(gr.abs().max().item(), 1e-7) + atol
        if not ok:
            bad.append(f"{n}: norm {gn:.6e} vs {ref:.6e}; samples {samp} vs {g['grad_samples'][i]}")
            wide = abs(gn - ref) <= 0.3 * ref + 10 * atol and \
                np.abs(samp - g["grad_samples"][i]).max() <= 0.3 * max(gr.abs().max().item(), 1e-7) + atol
            if not wide:
                very_bad.append(bad[-1])
    report(f"[parity] {tag} train gemm={gemm}: logits {lerr:.2e}, worst grad-norm rel err {worst['other']:.2e} "
            f"(attention-logit layers {worst['res']:.2e}), {len(bad)} of {len(names)} tensors outside the band")
    assert not very_bad, f"{len(very_bad)} of {len(names)} tensors\n" + "\n".join(very_bad[:60])
    assert len(bad) <= (0 if exact else 2), f"{len(bad)} of {len(names)} tensors\n" + "\n".join(bad[:60])
    for i, n in enumerate(str(s) for s in g["bn_names"]):
        bn = dict(model.named_modules())[n]
        assert np.abs(bn.running_mean[:4].cpu().numpy() - g["bn_rm4"][i]).max() < 1e-4, n
        assert np.abs(bn.running_var[:4].cpu().numpy() - g["bn_rv4"][i]).max() < 1e-4, n
