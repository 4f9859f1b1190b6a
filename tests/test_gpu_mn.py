"""GPU parity of the MN forward (eval) through the C ABI against the reference's golden vectors and the
oracle.  fp32 mode: logits within 1e-3 max-abs, top-10 indices identical (near-ties < 1e-4 apart in the
reference may swap).  bf16 mode: activations stored in bf16 -- tolerance 6e-2 stated here, reported in DESIGN.md."""
import numpy as np
import pytest
import torch

from oracle import net_oracle
from tests.util import report, NETS, build_model, fmap_digest, golden, net_inputs, topk_match

pytestmark = pytest.mark.gpu
MN_TAGS = ["mn10", "mn04", "mn20", "mn10_10s", "mn40_10s"]     # *_10s: the benchmarked 1000-frame shape


def _layer_report(tag, model, spec):
    """per-layer max-abs error against the oracle -- printed when a parity assertion fails."""
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        _, _, ref = net_oracle.mn_forward(sd, spec, width_mult=NETS[tag][1], return_fmaps=True)
        _, got = model._forward_impl(spec.cuda(), return_fmaps=True)
    lines = []
    for i, (r, g) in enumerate(zip(ref, got)):
        g = g.float().cpu()
        lines.append(f"fmap {i:2d} shape {tuple(r.shape)} max-abs err {(g - r).abs().max().item():.3e} "
                     f"(ref abs-max {r.abs().max().item():.3e})")
    return "\n".join(lines)


@pytest.mark.parametrize("gemm", ["simt", "auto"])
@pytest.mark.parametrize("tag", MN_TAGS)
def test_mn_eval_fp32_matches_reference_vectors(tag, gemm):
    """gemm = 'simt': exact fp32 CUDA-core GEMMs; 'auto': tcgen05 GEMMs (three-bf16-MMA fp32 emulation)."""
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
    assert np.abs(feat - g["eval_feat"]).max() < 1e-3
    assert topk_match(logits, g["eval_logits"], 10, tie_tol=1e-4)


@pytest.mark.parametrize("tag", ["mn10"])
def test_mn_eval_fmaps_match_oracle(tag):
    g = golden(tag)
    model = build_model(tag).cuda().eval()
    spec, _ = net_inputs(tag)
    with torch.no_grad():
        _, fmaps = model._forward_impl(spec.cuda(), return_fmaps=True)
    d = fmap_digest([f.float().cpu().contiguous() for f in fmaps])
    assert d.shape == g["eval_fmaps"].shape
    assert np.abs(d - g["eval_fmaps"]).max() < 2e-3, _layer_report(tag, model, spec)


@pytest.mark.parametrize("tag", ["mn10"])
def test_mn_eval_bf16_close(tag):
    g = golden(tag)
    model = build_model(tag, precision="bf16").cuda().eval()
    spec, _ = net_inputs(tag)
    with torch.no_grad():
        logits, _ = model(spec.cuda())
    err = np.abs(logits.cpu().numpy() - g["eval_logits"]).max()
    assert err < 6e-2, err


def test_mn_eval_batch1_and_odd_length():
    """inference.py path: B == 1, arbitrary T (T = 63 -> 32 style floor sizes)."""
    model = build_model("mn10").cuda().eval()
    spec = net_inputs("mn10")[0][:1, :, :, :173].contiguous()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        logits, feat = model(spec.cuda())
        ref_logits, ref_feat = net_oracle.mn_forward(sd, spec)
    assert logits.shape == (1, 527) and feat.shape == (1, 960)
    assert (logits.cpu() - ref_logits).abs().max() < 1e-3


def test_mn_rejects_cpu_input():
    model = build_model("mn04").eval()
    with pytest.raises(RuntimeError):
        model(torch.zeros(1, 1, 128, 100))
