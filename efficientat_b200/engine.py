"""Execution engine: walks an MN module tree and launches the fused sm_100a kernels.

The module tree (efficientat_b200.models.mn) only *holds* parameters; this file is the forward /
backward pass.  Activations are NHWC ([B, F, T, C]) in fp32 or bf16; parameters are read in place
from the nn.Parameters (fp32) on every call, so optimiser steps and load_state_dict need no cache
invalidation.  Everything is enqueued on the current CUDA stream through the C ABI
(include/eat_b200.h); no torch compute op is on the path (torch supplies allocation only).

Data flow per InvertedResidual (reference models/mn/block_types.py:177-181):
  eval :  [pw-GEMM +BN+act] -> dw conv +BN+act (+SE squeeze) -> [SE MLP] -> pw-GEMM (SE gate on load)
          +BN (+residual)                       -- folded BatchNorm, 3-4 launches per block
  train:  conv kernels emit RAW outputs + per-channel batch statistics; the BatchNorm+activation
          of layer l is applied on the operand load of layer l+1, so no normalised tensor is written
          except the block output (BN3 + residual).
"""
import os

import torch

from ._lib import lib

ACT = {"none": 0, "relu": 1, "hswish": 2}
BN_EPS = 1e-3
BN_MOM = 0.01


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _conv_out(n, k, s):
    return (n + 2 * ((k - 1) // 2) - k) // s + 1


class _Layer:
    """Plain record describing one block of the network for the launcher loops."""
    pass


class MNEngine:
    def __init__(self, model):
        self.model = model
        self.precision = getattr(model, "precision", "fp32")
        if self.precision not in ("fp32", "bf16"):
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {self.precision}")
        self.gemm_impl = os.environ.get("EAT_GEMM", "auto")     # auto | simt | tc
        self._plan()

    # ------------------------------------------------------------------ structure
    def _plan(self):
        from .models.mn.block_types import ConcurrentSEBlock, ConvNormActivation, InvertedResidual
        feats = list(self.model.features)
        self.stem = feats[0]
        self.last = feats[-1]
        self.blocks = []
        for m in feats[1:-1]:
            assert isinstance(m, InvertedResidual)
            L = _Layer()
            subs = list(m.block)
            L.expand = L.se = None
            i = 0
            if len([s for s in subs if isinstance(s, ConvNormActivation)]) == 3:
                L.expand = subs[0]
                i = 1
            L.dw = subs[i]
            i += 1
            if isinstance(subs[i], ConcurrentSEBlock):
                L.se = subs[i].conc_se_layers[0]
                i += 1
            L.proj = subs[i]
            L.res = m.use_res_connect
            L.act = ACT["hswish"] if m.cnf.use_hs else ACT["relu"]
            L.k = m.cnf.kernel
            L.stride = m.cnf.stride
            L.cin, L.cexp, L.cout = m.cnf.input_channels, m.cnf.expanded_channels, m.cnf.out_channels
            self.blocks.append(L)
        self.fc1 = self.model.classifier[2]
        self.fc2 = self.model.classifier[5]
        self.dropout_p = self.model.classifier[4].p

    @property
    def tdtype(self):
        return torch.float32 if self.precision == "fp32" else torch.bfloat16

    @property
    def dcode(self):
        return 0 if self.precision == "fp32" else 1

    # ------------------------------------------------------------------ kernel wrappers
    def _gemm(self, a, w, out, M, N, K, in_sc=None, in_act=0, gate=None, rows_per_sample=1, sc=None, bias=None,
              act=0, res=None, stats=None, a_code=None, c_code=None):
        """out[M,N] = epi(xf(a)[M,K] . w[N,K]^T).  in_sc: [2,K] (scale, shift) or None; sc: [2,N] or None;
        bias: [N] used as shift with scale None."""
        a_code = self.dcode if a_code is None else a_code
        c_code = self.dcode if c_code is None else c_code
        scale = _ptr(sc[0]) if sc is not None else 0
        shift = _ptr(sc[1]) if sc is not None else _ptr(bias)
        args = (a.data_ptr(), a_code, w.data_ptr(), out.data_ptr(), c_code, M, N, K,
                _ptr(in_sc[0]) if in_sc is not None else 0, _ptr(in_sc[1]) if in_sc is not None else 0, in_act,
                _ptr(gate), rows_per_sample, scale, shift, act, _ptr(res),
                _ptr(stats[0]) if stats is not None else 0, _ptr(stats[1]) if stats is not None else 0, _stream())
        L = lib()
        use_tc = self.gemm_impl == "tc" or (self.gemm_impl == "auto" and hasattr(L, "pw_tc_fwd") and
                                            a_code == c_code and M >= 128 and K % 8 == 0 and N % 8 == 0)
        if use_tc and hasattr(L, "pw_tc_fwd"):
            L.pw_tc_fwd(*args)
        else:
            L.gemm_simt_fwd(*args)

    def _fold(self, bn, dev):
        c = bn.num_features
        sc = torch.empty(2, c, device=dev, dtype=torch.float32)
        lib().bn_fold(bn.weight.data_ptr(), bn.bias.data_ptr(), bn.running_mean.data_ptr(),
                      bn.running_var.data_ptr(), bn.eps, sc[0].data_ptr(), sc[1].data_ptr(), c, _stream())
        return sc

    def _finalize(self, bn, stats, count, dev):
        """batch statistics -> (scale/shift [2,C], saved mean/invstd [2,C]); updates running buffers."""
        c = bn.num_features
        sc = torch.empty(2, c, device=dev, dtype=torch.float32)
        sv = torch.empty(2, c, device=dev, dtype=torch.float32)
        mom = bn.momentum if bn.momentum is not None else 0.1
        track = bn.track_running_stats and bn.running_mean is not None
        lib().bn_finalize(stats[0].data_ptr(), stats[1].data_ptr(), float(count), bn.weight.data_ptr(),
                          bn.bias.data_ptr(), bn.eps, mom, _ptr(bn.running_mean) if track else 0,
                          _ptr(bn.running_var) if track else 0, _ptr(bn.num_batches_tracked) if track else 0,
                          sc[0].data_ptr(), sc[1].data_ptr(), sv[0].data_ptr(), sv[1].data_ptr(), c, _stream())
        return sc, sv

    def _dw_weights(self, conv, dev):
        c, k = conv.out_channels, conv.kernel_size[0]
        wt = torch.empty(k * k, c, device=dev, dtype=torch.float32)
        lib().dw_repack(conv.weight.data_ptr(), wt.data_ptr(), c, k, _stream())
        return wt

    # ------------------------------------------------------------------ forward
    def forward(self, x, return_fmaps=False):
        if not x.is_cuda:
            raise RuntimeError("efficientat_b200 models run on CUDA (sm_100a) only; got a CPU tensor")
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape [B, 1, F, T], got {tuple(x.shape)}")
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters())
        if self.model.training:
            if return_fmaps:
                raise NotImplementedError("return_fmaps is available in eval mode only")
            from .autograd import mn_train_forward
            logits, feat = mn_train_forward(self, x, needs_grad)
            return self._squeeze(logits, feat) + (None,)
        if needs_grad and x.requires_grad:
            raise NotImplementedError("gradients w.r.t. the input spectrogram are not implemented")
        logits, feat, fmaps = self._forward_eval(x.detach(), return_fmaps)
        return self._squeeze(logits, feat) + (fmaps,)

    @staticmethod
    def _squeeze(logits, feat):
        # reference: `.squeeze()` then re-add the batch dim when B == 1 (mn/model.py:220-226); for B > 1
        # and num_classes > 1 the squeeze is a no-op, for B == 1 the unsqueeze restores [1, C].
        return logits, feat

    def _forward_eval(self, x, return_fmaps=False):
        L = lib()
        dev = x.device
        st = _stream()
        td, dc = self.tdtype, self.dcode
        x = x.float().contiguous()
        B, _, F, T = x.shape
        fmaps = [] if return_fmaps else None

        def keep(t, f, tt, c):
            if fmaps is not None:   # NHWC storage -> logical NCHW view, like the reference's fmaps
                fmaps.append(t.view(B, f, tt, c).permute(0, 3, 1, 2))

        conv, bn = self.stem[0], self.stem[1]
        s0 = conv.stride[0]
        Fi, Ti = _conv_out(F, 3, s0), _conv_out(T, 3, s0)
        c0 = conv.out_channels
        a = torch.empty(B, Fi, Ti, c0, device=dev, dtype=td)
        sc = self._fold(bn, dev)
        L.stem_fwd(x.data_ptr(), conv.weight.data_ptr(), a.data_ptr(), dc, B, F, T, c0, s0, sc[0].data_ptr(),
                   sc[1].data_ptr(), ACT["hswish"], 0, 0, st)
        keep(a, Fi, Ti, c0)
        for blk in self.blocks:
            inp = a
            M = B * Fi * Ti
            if blk.expand is not None:
                e = torch.empty(B, Fi, Ti, blk.cexp, device=dev, dtype=td)
                self._gemm(inp, blk.expand[0].weight, e, M, blk.cexp, blk.cin, sc=self._fold(blk.expand[1], dev),
                           act=blk.act)
            else:
                e = inp
            Fo, To = _conv_out(Fi, blk.k, blk.stride), _conv_out(Ti, blk.k, blk.stride)
            d = torch.empty(B, Fo, To, blk.cexp, device=dev, dtype=td)
            sc = self._fold(blk.dw[1], dev)
            pool = torch.zeros(B, blk.cexp, device=dev, dtype=torch.float32) if blk.se is not None else None
            L.dw_conv_fwd(e.data_ptr(), self._dw_weights(blk.dw[0], dev).data_ptr(), d.data_ptr(), dc, B, Fi, Ti,
                          blk.cexp, blk.k, blk.stride, 0, 0, 0, sc[0].data_ptr(), sc[1].data_ptr(), blk.act,
                          _ptr(pool), 0, 0, st)
            gate = None
            if blk.se is not None:
                gate = torch.empty(B, blk.cexp, device=dev, dtype=torch.float32)
                S = blk.se.fc1.out_features
                L.se_fc_fwd(pool.data_ptr(), 1.0 / (Fo * To), blk.se.fc1.weight.data_ptr(),
                            blk.se.fc1.bias.data_ptr(), blk.se.fc2.weight.data_ptr(), blk.se.fc2.bias.data_ptr(),
                            gate.data_ptr(), 0, B, blk.cexp, S, st)
            Mo = B * Fo * To
            o = torch.empty(B, Fo, To, blk.cout, device=dev, dtype=td)
            self._gemm(d, blk.proj[0].weight, o, Mo, blk.cout, blk.cexp, gate=gate, rows_per_sample=Fo * To,
                       sc=self._fold(blk.proj[1], dev), act=0, res=inp if blk.res else None)
            a, Fi, Ti = o, Fo, To
            keep(a, Fi, Ti, blk.cout)
        conv, bn = self.last[0], self.last[1]
        cl = conv.out_channels
        M = B * Fi * Ti
        z = torch.empty(B, Fi, Ti, cl, device=dev, dtype=td)
        self._gemm(a, conv.weight, z, M, cl, conv.in_channels, sc=self._fold(bn, dev), act=ACT["hswish"])
        keep(z, Fi, Ti, cl)
        feat = torch.zeros(B, cl, device=dev, dtype=torch.float32)
        L.bn_act_pool(z.data_ptr(), 0, 0, 0, feat.data_ptr(), 1.0 / (Fi * Ti), dc, B, Fi * Ti, cl, st)
        h = torch.empty(B, self.fc1.out_features, device=dev, dtype=torch.float32)
        self._gemm(feat, self.fc1.weight, h, B, self.fc1.out_features, cl, bias=self.fc1.bias, act=ACT["hswish"],
                   a_code=0, c_code=0)
        logits = torch.empty(B, self.fc2.out_features, device=dev, dtype=torch.float32)
        self._gemm(h, self.fc2.weight, logits, B, self.fc2.out_features, self.fc1.out_features, bias=self.fc2.bias,
                   a_code=0, c_code=0)
        return logits, feat, fmaps
