"""Pin the CPU oracle (oracle/) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle, net_oracle
from tests.util import GOLDEN, NETS, build_model, fmap_digest, golden, net_inputs, topk_match
from efficientat_b200.synth import synth_waveform

torch.set_num_threads(8)


def test_mel_oracle_matches_reference_vectors():
    g = golden("mel")
    y = mel_oracle.mel_forward(synth_waveform(2, 32000, seed=11))
    assert y.shape == g["y_noise"].shape
    assert np.abs(y.numpy() - g["y_noise"]).max() < 1e-4
    y2 = mel_oracle.mel_forward(synth_waveform(1, 5000, seed=12))
    assert y2.shape == g["y_ragged"].shape == (1, 128, 16)
    assert np.abs(y2.numpy() - g["y_ragged"]).max() < 1e-4
    y3 = mel_oracle.mel_forward(synth_waveform(1, 16000, seed=13), n_mels=64, hopsize=500, fmin=50.0, fmax=14000.0)
    assert np.abs(y3.numpy() - g["y_geom"]).max() < 1e-4
    # coloured signal: fp32 FFT noise floors differ between FFT implementations in near-silent bins, so
    # compare against the fp64 evaluation of the oracle with a tolerance on the log-mel
    y4 = mel_oracle.mel_forward(torch.from_numpy(g["x_tone"]), dtype=torch.float64)
    assert np.abs(y4.numpy() - g["y_tone"]).max() < 2e-3


def test_mel_oracle_fp64_close_to_fp32():
    x = synth_waveform(1, 32000, seed=3)
    a = mel_oracle.mel_forward(x, dtype=torch.float32)
    b = mel_oracle.mel_forward(x, dtype=torch.float64)
    assert (a.double() - b).abs().max() < 1e-4


@pytest.mark.parametrize("tag", list(NETS))
def test_state_dict_contract(tag):
    """module tree of this package has exactly the reference's state_dict keys and shapes (App. D)."""
    want = json.load(open(os.path.join(GOLDEN, f"statedict_{tag.split('_')[0]}.json")))
    got = {k: list(v.shape) for k, v in build_model(tag).state_dict().items()}
    assert list(got) == list(want)
    assert got == want


def _oracle_forward(tag, sd, spec, **kw):
    kind, width, _, _ = NETS[tag]
    fn = net_oracle.mn_forward if kind == "mn" else net_oracle.dymn_forward
    if kind == "dymn":
        kw.setdefault("temperature", 30.0)      # get_model default T_max without pretrained weights
    return fn(sd, spec, width_mult=width, **kw)


@pytest.mark.parametrize("tag", list(NETS))
def test_net_oracle_eval_matches_reference_vectors(tag):
    g = golden(tag)
    model = build_model(tag)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    spec, _ = net_inputs(tag)
    assert np.abs(fmap_digest([spec]) - g["spec_digest"]).max() == 0
    with torch.no_grad():
        logits, feat, fmaps = _oracle_forward(tag, sd, spec, return_fmaps=True)
    assert np.abs(logits.numpy() - g["eval_logits"]).max() < 1e-4
    assert np.abs(feat.numpy() - g["eval_feat"]).max() < 2e-4
    d = fmap_digest(fmaps)
    assert d.shape == g["eval_fmaps"].shape
    assert np.abs(d - g["eval_fmaps"]).max() < 1e-3 * max(1.0, np.abs(g["eval_fmaps"]).max())
    assert topk_match(logits.numpy(), g["eval_logits"], 10, tie_tol=1e-5)


@pytest.mark.parametrize("tag", ["mn10", "mn04", "dymn10", "dymn04", "mn10_10s", "dymn20"])
def test_net_oracle_train_matches_reference_vectors(tag):
    g = golden(tag)
    model = build_model(tag)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    params = {k for k, _ in model.named_parameters()}
    for k in params:
        sd[k].requires_grad_(True)
    spec, y = net_inputs(tag)
    logits, _ = _oracle_forward(tag, sd, spec, training=True)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y)
    loss.backward()
    assert abs(loss.item() - float(g["train_loss"])) < 1e-5
    assert np.abs(logits.detach().numpy() - g["train_logits"]).max() < 1e-4
    names = [str(n) for n in g["grad_names"]]
    assert set(names) == params
    for i, n in enumerate(names):
        gn = sd[n].grad.double().norm().item()
        assert abs(gn - g["grad_norm"][i]) <= 2e-3 * g["grad_norm"][i] + 1e-7, n
    for i, n in enumerate(str(s) for s in g["bn_names"]):
        assert np.abs(sd[n + ".running_mean"][:4].numpy() - g["bn_rm4"][i]).max() < 1e-5, n
        assert np.abs(sd[n + ".running_var"][:4].numpy() - g["bn_rv4"][i]).max() < 1e-5, n
