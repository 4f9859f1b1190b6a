"""Deterministic synthetic model state and inputs.

Default-initialised reference models emit logits of ~1e-10 (Linear ~ N(0, 0.01), BN identity),
which makes a "max-abs <= 1e-3" parity check vacuous.  ``synth_state_`` rewrites every
parameter/buffer of a model *in place* from per-tensor seeded generators, so the same
recipe applied to the reference modules, to the oracle's state_dict and to this package's
modules yields bit-identical tensors regardless of construction order.
Works by duck-typing on module class names so that it can be pointed at the reference's
own modules (tests/golden/make_golden.py) without importing them here.
"""
import zlib

import torch


CONV_GAIN = 1.0


def _gen(seed, name):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def _fill(t, kind, g, a=0.0, b=1.0):
    tmp = torch.empty(t.shape, dtype=torch.float32)
    if kind == "normal":
        tmp.normal_(a, b, generator=g)
    else:
        tmp.uniform_(a, b, generator=g)
    with torch.no_grad():
        t.copy_(tmp.to(t.dtype))


@torch.no_grad()
def synth_state_(model, seed=0):
    """BN: running_mean~N(0,.1), running_var~U(.5,1.5), weight~U(.5,1.5), bias~N(0,.1);
    Linear: weight~N(0,1/fan_in), bias~N(0,.1); Conv2d and DynamicConv banks: N(0, gain/fan_in) with
    fan_in = (Cin/groups)*k*k (+bias N(0,.1)).  fan_in scaling (not the reference's kaiming fan_out
    init, mn/model.py:201) is deliberate: with fan_out a depthwise layer shrinks the signal by ~1/C and,
    BatchNorm being frozen in eval mode, the logits stop depending on the input after a few blocks --
    a parity check on such a state would be blind to errors in the early layers."""
    for name, m in model.named_modules():
        cls = type(m).__name__
        if cls == "BatchNorm2d":
            _fill(m.running_mean, "normal", _gen(seed, name + ".running_mean"), 0.0, 0.1)
            _fill(m.running_var, "uniform", _gen(seed, name + ".running_var"), 0.5, 1.5)
            _fill(m.weight, "uniform", _gen(seed, name + ".weight"), 0.5, 1.5)
            _fill(m.bias, "normal", _gen(seed, name + ".bias"), 0.0, 0.1)
            m.num_batches_tracked.zero_()
        elif cls == "Linear":
            _fill(m.weight, "normal", _gen(seed, name + ".weight"), 0.0, (1.0 / m.in_features) ** 0.5)
            if m.bias is not None:
                _fill(m.bias, "normal", _gen(seed, name + ".bias"), 0.0, 0.1)
        elif cls == "Conv2d":
            fan_in = (m.in_channels // m.groups) * m.kernel_size[0] * m.kernel_size[1]
            _fill(m.weight, "normal", _gen(seed, name + ".weight"), 0.0, (CONV_GAIN / fan_in) ** 0.5)
            if m.bias is not None:
                _fill(m.bias, "normal", _gen(seed, name + ".bias"), 0.0, 0.1)
        elif cls == "DynamicConv":
            fan_in = (m.in_channels // m.groups) * m.kernel_size * m.kernel_size
            _fill(m.weight, "normal", _gen(seed, name + ".weight"), 0.0, (CONV_GAIN / fan_in) ** 0.5)
    return model


def synth_waveform(batch, n_samples=320000, seed=0, std=0.1):
    """[B, N] fp32 CPU waveform ~ N(0, std^2), one generator per clip so that clip i does not
    depend on the batch size."""
    out = torch.empty(batch, n_samples, dtype=torch.float32)
    for i in range(batch):
        out[i].normal_(0.0, std, generator=_gen(seed, f"wave{i}"))
    return out


def synth_labels(batch, n_classes=527, seed=0, p=0.005):
    g = _gen(seed, "labels")
    return (torch.rand(batch, n_classes, generator=g) < p).float()


def bn_modules(model):
    return [(n, m) for n, m in model.named_modules() if type(m).__name__ == "BatchNorm2d"]


@torch.no_grad()
def get_bn_stats(model):
    """-> (running_mean, running_var) of every BatchNorm2d concatenated in module order."""
    ms = bn_modules(model)
    return (torch.cat([m.running_mean.flatten().float().cpu() for _, m in ms]),
            torch.cat([m.running_var.flatten().float().cpu() for _, m in ms]))


@torch.no_grad()
def set_bn_stats(model, rm, rv):
    """Inverse of get_bn_stats: used to install *calibrated* running statistics (statistics of a
    real forward pass), without which a randomly initialised eval-mode network saturates or dies."""
    off = 0
    for _, m in bn_modules(model):
        c = m.num_features
        m.running_mean.copy_(torch.as_tensor(rm[off:off + c]).to(m.running_mean))
        m.running_var.copy_(torch.as_tensor(rv[off:off + c]).to(m.running_var))
        off += c
    assert off == len(rm)
    return model
