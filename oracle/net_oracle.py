"""Oracle (test infrastructure, not product code): MN and DyMN networks.

Functional restatement -- state_dict in, tensors out -- of
  MN            reference models/mn/model.py:73-234, models/mn/block_types.py:45-181
  DyMN          reference models/dymn/model.py:36-206, models/dymn/dy_block.py:44-409
built from stock torch.nn.functional ops in NCHW.  Works in fp32 or fp64, on CPU
(and on CUDA, where it doubles as the stock cuDNN-path timing arm).
Training mode uses per-batch BatchNorm statistics and updates running buffers in
place, like nn.BatchNorm2d(eps=1e-3, momentum=0.01) (mn/model.py:114-115,
dy_block.py:284).
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-3
BN_MOM = 0.01

# (in, kernel, expanded, out, use_se, act, stride-slot)  -- mn/model.py:252-268, dymn/model.py:233-251
_ROWS = [
    (16, 3, 16, 16, False, "RE", None),
    (16, 3, 64, 24, False, "RE", 0),
    (24, 3, 72, 24, False, "RE", None),
    (24, 5, 72, 40, True, "RE", 1),
    (40, 5, 120, 40, True, "RE", None),
    (40, 5, 120, 40, True, "RE", None),
    (40, 3, 240, 80, False, "HS", 2),
    (80, 3, 200, 80, False, "HS", None),
    (80, 3, 184, 80, False, "HS", None),
    (80, 3, 184, 80, False, "HS", None),
    (80, 3, 480, 112, True, "HS", None),
    (112, 3, 672, 112, True, "HS", None),
    (112, 5, 672, 160, True, "HS", 3),
    (160, 5, 960, 160, True, "HS", None),
    (160, 5, 960, 160, True, "HS", None),
]


def make_divisible(v, divisor=8):
    """mn/utils.py:8-21"""
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def block_table(width_mult=1.0, strides=(2, 2, 2, 2)):
    rows = []
    for cin, k, cexp, cout, se, act, slot in _ROWS:
        rows.append(dict(cin=make_divisible(cin * width_mult), k=k, cexp=make_divisible(cexp * width_mult),
                         cout=make_divisible(cout * width_mult), se=se, act=act,
                         stride=1 if slot is None else strides[slot]))
    return rows


def _act(x, kind):
    if kind == "HS":
        return F.hardswish(x)
    if kind == "RE":
        return F.relu(x)
    return x


def _bn(x, sd, pfx, training):
    rm, rv = sd[pfx + ".running_mean"], sd[pfx + ".running_var"]
    if training and (pfx + ".num_batches_tracked") in sd:
        sd[pfx + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, sd[pfx + ".weight"], sd[pfx + ".bias"], training, BN_MOM, BN_EPS)


def _cna(x, sd, pfx, k, stride, groups, act, training):
    """torchvision ConvNormActivation: conv(bias=False, padding=(k-1)//2) -> BN -> act."""
    x = F.conv2d(x, sd[pfx + ".0.weight"], None, stride, (k - 1) // 2, 1, groups)
    return _act(_bn(x, sd, pfx + ".1", training), act)


def _se(x, sd, pfx):
    """block_types.py:72-83 with se_dims='c' (mean over F,T; ReLU; Sigmoid gate)."""
    s = x.mean((2, 3))
    s = F.relu(F.linear(s, sd[pfx + ".fc1.weight"], sd[pfx + ".fc1.bias"]))
    s = torch.sigmoid(F.linear(s, sd[pfx + ".fc2.weight"], sd[pfx + ".fc2.bias"]))
    return x * s[:, :, None, None]


def _head(x, sd, training, dropout_mask=None):
    """mn/model.py:187-194,220-221: avgpool -> Linear -> Hardswish -> Dropout(0.2) -> Linear."""
    feat = x.mean((2, 3))
    h = F.hardswish(F.linear(feat, sd["classifier.2.weight"], sd["classifier.2.bias"]))
    if training and dropout_mask is not None:
        h = h * dropout_mask
    return F.linear(h, sd["classifier.5.weight"], sd["classifier.5.bias"]), feat


def mn_forward(sd, x, width_mult=1.0, strides=(2, 2, 2, 2), training=False, dropout_mask=None,
               return_fmaps=False):
    """x: [B,1,F,T] -> (logits [B,classes], features [B,960w]).  dropout_mask: optional
    [B,1280w] multiplier (already scaled by 1/(1-p)) standing in for nn.Dropout in training."""
    rows = block_table(width_mult, strides)
    fmaps = []
    x = _cna(x, sd, "features.0", 3, 2, 1, "HS", training)
    fmaps.append(x)
    for i, r in enumerate(rows, start=1):
        inp = x
        j = 0
        pfx = f"features.{i}.block"
        if r["cexp"] != r["cin"]:
            x = _cna(x, sd, f"{pfx}.{j}", 1, 1, 1, r["act"], training)
            j += 1
        x = _cna(x, sd, f"{pfx}.{j}", r["k"], r["stride"], r["cexp"], r["act"], training)
        j += 1
        if r["se"]:
            x = _se(x, sd, f"{pfx}.{j}.conc_se_layers.0")
            j += 1
        x = _cna(x, sd, f"{pfx}.{j}", 1, 1, 1, None, training)
        if r["stride"] == 1 and r["cin"] == r["cout"]:
            x = x + inp
        fmaps.append(x)
    x = _cna(x, sd, f"features.{len(rows) + 1}", 1, 1, 1, "HS", training)
    fmaps.append(x)
    logits, feat = _head(x, sd, training, dropout_mask)
    if return_fmaps:
        return logits, feat, fmaps
    return logits, feat


# ----------------------------------------------------------------------------- DyMN

def dymn_context_dim(cexp, width_mult, context_ratio=4, min_ctx=32, max_ctx=128):
    """dy_block.py:278-281"""
    lo, hi = make_divisible(min_ctx * width_mult), make_divisible(max_ctx * width_mult)
    return int(min(max(make_divisible(cexp // context_ratio), lo), hi))


def _dyconv(x, h_c, sd, pfx, cout, cin_g, k, stride, groups, temperature):
    """dy_block.py:103-131: softmax(Linear(h_c)/T) mixes K kernels; per-sample grouped conv."""
    b = x.shape[0]
    att = F.softmax(F.linear(h_c, sd[pfx + ".residuals.0.weight"], sd[pfx + ".residuals.0.bias"]) / temperature, -1)
    w = att @ sd[pfx + ".weight"][0, 0]                      # [B, cout*cin_g*k*k]
    w = w.reshape(b * cout, cin_g, k, k)
    y = F.conv2d(x.reshape(1, -1, x.shape[2], x.shape[3]), w, None, stride, (k - 1) // 2, 1, groups * b)
    return y.reshape(b, cout, y.shape[2], y.shape[3])


def _context_gen(x, sd, pfx, stride, training):
    """dy_block.py:235-254"""
    cf = x.mean(3, keepdim=True)                              # [B,C,F,1]
    ct = x.mean(2, keepdim=True).permute(0, 1, 3, 2)          # [B,C,T,1]
    f = cf.shape[2]
    g = torch.cat([cf, ct], 2)
    g = F.conv2d(g, sd[pfx + ".joint_conv.weight"])
    g = F.hardswish(_bn(g, sd, pfx + ".joint_norm", training))
    h_cf, h_ct = g[:, :, :f], g[:, :, f:].permute(0, 1, 3, 2)
    h_c = g.mean(2).flatten(1)                                # [B,H]
    if stride > 1:
        h_cf = F.avg_pool2d(h_cf, (3, 1), (stride, 1), (1, 0))
        h_ct = F.avg_pool2d(h_ct, (1, 3), (1, stride), (0, 1))
    g_cf = F.conv2d(h_cf, sd[pfx + ".conv_f.weight"], sd[pfx + ".conv_f.bias"])
    g_ct = F.conv2d(h_ct, sd[pfx + ".conv_t.weight"], sd[pfx + ".conv_t.bias"])
    return h_c, g_cf, g_ct


def _dyrelu_b(x, h_c, sd, pfx):
    """dy_block.py:157-188 with M=2: max(a1*x+b1, a2*x+b2), coefs = (2*sigmoid(.)-1)*lambdas+init_v."""
    b, c = x.shape[:2]
    theta = 2 * torch.sigmoid(F.linear(h_c, sd[pfx + ".coef_net.0.weight"], sd[pfx + ".coef_net.0.bias"])) - 1
    co = theta.view(b, c, 4) * sd[pfx + ".lambdas"] + sd[pfx + ".init_v"]
    a1, a2, b1, b2 = (co[:, :, i, None, None] for i in range(4))
    return torch.maximum(x * a1 + b1, x * a2 + b2)


def dymn_forward(sd, x, width_mult=1.0, strides=(2, 2, 2, 2), training=False, temperature=1.0,
                 dropout_mask=None, return_fmaps=False):
    """All-dynamic DyMN (use_dy_blocks='all').  dymn/model.py:157-200, dy_block.py:390-409."""
    rows = block_table(width_mult, strides)
    fmaps = []
    x = _cna(x, sd, "in_c", 3, 2, 1, "HS", training)
    fmaps.append(x)
    for i, r in enumerate(rows):
        inp = x
        p = f"layers.{i}"
        h_c, g_cf, g_ct = _context_gen(x, sd, p + ".context_gen", r["stride"], training)
        if r["cexp"] != r["cin"]:
            x = _dyconv(x, h_c, sd, p + ".exp_conv", r["cexp"], r["cin"], 1, 1, 1, temperature)
            x = _act(_bn(x, sd, p + ".exp_norm", training), r["act"])
        x = _dyconv(x, h_c, sd, p + ".depth_conv", r["cexp"], 1, r["k"], r["stride"], r["cexp"], temperature)
        x = _bn(x, sd, p + ".depth_norm", training)
        x = _dyrelu_b(x, h_c, sd, p + ".depth_act")
        x = x * torch.sigmoid(g_cf) * torch.sigmoid(g_ct)     # CoordAtt dy_block.py:195-201
        x = _dyconv(x, h_c, sd, p + ".proj_conv", r["cout"], r["cexp"], 1, 1, 1, temperature)
        x = _bn(x, sd, p + ".proj_norm", training)
        if r["stride"] == 1 and r["cin"] == r["cout"]:
            x = x + inp
        fmaps.append(x)
    x = _cna(x, sd, "out_c", 1, 1, 1, "HS", training)
    fmaps.append(x)
    logits, feat = _head(x, sd, training, dropout_mask)
    if return_fmaps:
        return logits, feat, fmaps
    return logits, feat


def dyconv_temperature(epoch, t_max=30.0, t_min=1.0, t0_slope=1.0, t1_slope=0.02):
    """dy_block.py:133-139"""
    t0 = t_max - t0_slope * epoch
    t1 = 1 + t1_slope * (t_max - 1) / t0_slope - t1_slope * epoch
    return max(t0, t1, t_min)
