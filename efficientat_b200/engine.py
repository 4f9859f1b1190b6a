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

from ._lib import check_module_tensors, lib

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


# measured (scripts/gpu_runs/r2_sefuse1.sh, mn10 B=256): 34.3 -> 33.76 ms/step; EAT_SE_FUSED=0 restores the two-pass route
SE_FUSED_DEFAULT = "1"


DGRAD_BNRED_DEFAULT = "0"


class _ZeroPool:
    """fp64 accumulators for BatchNorm statistics, carved from chunks that are zeroed with ONE fill each
    (a training step needs ~100 small zeroed buffers; one launch per buffer showed up as 250 tiny kernels)."""

    def __init__(self, chunk=1 << 16):
        self.chunk, self.buf, self.off = chunk, None, 0

    def take(self, rows, cols, dev):
        n = rows * cols
        if self.buf is None or self.off + n > self.buf.numel() or self.buf.device != dev:
            self.buf = torch.zeros(max(self.chunk, n), device=dev, dtype=torch.float64)
            self.off = 0
        v = self.buf[self.off:self.off + n].view(rows, cols)
        self.off += n
        return v


class _Fork:
    """Weight gradients are leaves of the backward graph: nothing downstream waits for them.  They are launched on a side
    stream so that they run beside the data-gradient / BatchNorm chain of the same block (in a captured CUDA graph the
    fork and join become graph edges): latency-bound tails of one kernel are filled by the other, and the two readers of
    a gradient tensor (weight- and data-gradient GEMM) sweep it together, so its second read tends to hit L2.
    Tensors the side stream reads are kept alive until `join`, after which the main stream may recycle them."""

    def __init__(self, device, enabled):
        self.enabled = enabled
        self.side = torch.cuda.Stream(device) if enabled else None
        self.keep = []

    def run(self, fn, *tensors):
        if not self.enabled:
            fn()
            return
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            fn()
        self.keep.extend(tensors)

    def join(self):
        if self.enabled:
            torch.cuda.current_stream().wait_stream(self.side)
            self.keep.clear()


class MNEngine:
    def __init__(self, model):
        self.model = model
        self.precision = getattr(model, "precision", "fp32")
        if self.precision not in ("fp32", "bf16"):
            raise ValueError(f"precision must be 'fp32' or 'bf16', got {self.precision}")
        self.gemm_impl = os.environ.get("EAT_GEMM", "auto")     # auto | simt (exact-fp32 CUDA cores everywhere)
        self.tc_min_rows = 1024                                  # tiny GEMMs (classifier, SE) stay on CUDA cores
        self.pw_impl = "tc" if os.environ.get("EAT_PW_IMPL") == "tc" else "tma"
        # weight gradients on a side stream (see _Fork): measured +1.1 % at B=256 (36.6 vs 37.0 ms/step,
        # profiles/README.md) -- both branches are HBM-bound -- so it stays an opt-in experiment
        self.fork_wgrad = os.environ.get("EAT_FORK_WGRAD", "0") == "1"
        # SE blocks: squeeze-excitation reduce + BatchNorm-backward reduce of the depthwise output in one pass over the two
        # expanded tensors (eat_se_bn_bwd_reduce / _combine) instead of two (eat_se_bwd_reduce, eat_bn_bwd_reduce)
        self.se_fused = os.environ.get("EAT_SE_FUSED", SE_FUSED_DEFAULT) == "1"
        self.se_parts_per_sm = int(os.environ.get("EAT_SE_PARTS_PER_SM", "6"))
        # stride-2 blocks: the BatchNorm-backward reduce of the expand stage inside the depthwise data-gradient kernel that
        # produces its upstream gradient (eat_dw_conv_dgrad_bnred), fp32 storage
        self.dgrad_bnred = os.environ.get("EAT_DGRAD_BNRED", DGRAD_BNRED_DEFAULT) == "1"
        self._fork = None
        self._se_scale = {}
        self._zero_pool = _ZeroPool()
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
              act=0, res=None, stats=None, a_code=None, c_code=None, w_trans=False):
        """out[M,N] = epi(xf(a)[M,K] . w[N,K]^T).  in_sc: [2,K] (scale, shift) or None; sc: [2,N] or None;
        bias: [N] used as shift with scale None."""
        a_code = self.dcode if a_code is None else a_code
        c_code = self.dcode if c_code is None else c_code
        scale = _ptr(sc[0]) if sc is not None else 0
        shift = _ptr(sc[1]) if sc is not None else _ptr(bias)
        args = (a.data_ptr(), a_code, w.data_ptr(), 1 if w_trans else 0, out.data_ptr(), c_code, M, N, K,
                _ptr(in_sc[0]) if in_sc is not None else 0, _ptr(in_sc[1]) if in_sc is not None else 0, in_act,
                _ptr(gate), rows_per_sample, scale, shift, act, _ptr(res),
                _ptr(stats[0]) if stats is not None else 0, _ptr(stats[1]) if stats is not None else 0, _stream())
        L = lib()
        use_tc = (self.gemm_impl != "simt" and a_code == c_code and M >= self.tc_min_rows and K % 8 == 0
                  and N % 8 == 0 and act != 3 and in_act != 3)      # sigmoid epilogues (DyMN context nets) stay on CUDA cores
        if use_tc and a_code == 0 and self.pw_impl == "tma" and not (res is not None and act != 0):
            # fp32 storage: TMA-fed kernel; the weights are pre-split (bf16 hi|lo rows, BN scale folded, transposed for the
            # data gradient) once per launch into this scratch, so no CTA repeats that per tile
            ws = torch.empty(N * ((K + 31) // 32) * 128, device=w.device, dtype=torch.uint8)
            L.pw_tma_fwd(a.data_ptr(), w.data_ptr(), 1 if w_trans else 0, out.data_ptr(), M, N, K, args[9], args[10], in_act,
                         _ptr(gate), rows_per_sample, scale, shift, act, _ptr(res), args[18], args[19], ws.data_ptr(),
                         ws.numel(), _stream())
        elif use_tc:
            if w_trans:      # data gradient: feed W^T [N, K] as a K-major operand
                wt = torch.empty(N, K, device=w.device, dtype=torch.float32)
                L.transpose_f32(w.data_ptr(), wt.data_ptr(), K, N, _stream())
                args = (args[0], args[1], wt.data_ptr(), 0) + args[4:]
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

    def _se_gate(self, se, pool, inv_count, B, C, dev, hidden=None):
        """Squeeze-excitation MLP (block_types.py:72-83) as two batched GEMMs: gate = sigmoid(W2 relu(W1 mean + b1) + b2)
        with mean = pool * inv_count folded into the first epilogue (scale = inv_count, shift = b1)."""
        S = se.fc1.out_features
        key = (S, float(inv_count), str(dev))
        sc = self._se_scale.get(key)
        if sc is None:
            sc = torch.full((S,), float(inv_count), device=dev, dtype=torch.float32)
            self._se_scale[key] = sc
        if hidden is None:
            hidden = torch.empty(B, S, device=dev, dtype=torch.float32)
        self._gemm(pool, se.fc1.weight, hidden, B, S, C, sc=(sc, se.fc1.bias), act=ACT["relu"], a_code=0, c_code=0)
        gate = torch.empty(B, C, device=dev, dtype=torch.float32)
        self._gemm(hidden, se.fc2.weight, gate, B, C, S, bias=se.fc2.bias, act=3, a_code=0, c_code=0)
        return gate, hidden

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
        check_module_tensors(self.model, x.device, type(self.model).__name__)
        self.dropout_p = float(self.model.classifier[4].p)          # read at call time (it may be changed after engine())
        needs_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters())
        with torch.cuda.device(x.device):                            # launches go to x's device, whatever is current
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
            a, Fi, Ti = self._ir_block_eval(blk, a, B, Fi, Ti)
            keep(a, Fi, Ti, blk.cout)
        logits, feat, z = self._head_eval(a, B, Fi, Ti)
        keep(z, Fi, Ti, self.last[0].out_channels)
        return logits, feat, fmaps

    def _ir_block_eval(self, blk, a, B, Fi, Ti):
        """one InvertedResidual with folded BatchNorm: 3-4 launches (reference block_types.py:177-181)"""
        L = lib()
        dev, st = a.device, _stream()
        td, dc = self.tdtype, self.dcode
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
            gate, _ = self._se_gate(blk.se, pool, 1.0 / (Fo * To), B, blk.cexp, dev)
        Mo = B * Fo * To
        o = torch.empty(B, Fo, To, blk.cout, device=dev, dtype=td)
        self._gemm(d, blk.proj[0].weight, o, Mo, blk.cout, blk.cexp, gate=gate, rows_per_sample=Fo * To,
                   sc=self._fold(blk.proj[1], dev), act=0, res=inp if blk.res else None)
        return o, Fo, To

    def _head_eval(self, a, B, Fi, Ti):
        """last 1x1 conv + BN + Hardswish, global average pool, classifier MLP (mn/model.py:160-166,187-194,220-221)"""
        L = lib()
        dev, st = a.device, _stream()
        td, dc = self.tdtype, self.dcode
        conv, bn = self.last[0], self.last[1]
        cl = conv.out_channels
        M = B * Fi * Ti
        z = torch.empty(B, Fi, Ti, cl, device=dev, dtype=td)
        self._gemm(a, conv.weight, z, M, cl, conv.in_channels, sc=self._fold(bn, dev), act=ACT["hswish"])
        feat = torch.zeros(B, cl, device=dev, dtype=torch.float32)
        L.bn_act_pool(z.data_ptr(), 0, 0, 0, feat.data_ptr(), 1.0 / (Fi * Ti), dc, B, Fi * Ti, cl, st)
        h = torch.empty(B, self.fc1.out_features, device=dev, dtype=torch.float32)
        self._gemm(feat, self.fc1.weight, h, B, self.fc1.out_features, cl, bias=self.fc1.bias, act=ACT["hswish"],
                   a_code=0, c_code=0)
        logits = torch.empty(B, self.fc2.out_features, device=dev, dtype=torch.float32)
        self._gemm(h, self.fc2.weight, logits, B, self.fc2.out_features, self.fc1.out_features, bias=self.fc2.bias,
                   a_code=0, c_code=0)
        return logits, feat, z

    # ------------------------------------------------------------------ training forward
    def _new_stats(self, c, dev):
        return self._zero_pool.take(2, c, dev)

    def _forward_train(self, x, dropout_mask=None):
        """Batch-statistics forward.  Returns (logits, feat, saved) where `saved` holds the raw conv outputs
        and BatchNorm statistics needed by `_backward` (activations are recomputed from them)."""
        L = lib()
        dev = x.device
        st = _stream()
        td, dc = self.tdtype, self.dcode
        x = x.detach().float().contiguous()
        B, _, F, T = x.shape
        HS = ACT["hswish"]
        S = {"x": x, "B": B, "F": F, "T": T, "blocks": []}
        self._zero_pool = _ZeroPool()

        conv, bn = self.stem[0], self.stem[1]
        s0 = conv.stride[0]
        Fi, Ti = _conv_out(F, 3, s0), _conv_out(T, 3, s0)
        c0 = conv.out_channels
        z0 = torch.empty(B, Fi, Ti, c0, device=dev, dtype=td)
        stt = self._new_stats(c0, dev)
        L.stem_fwd(x.data_ptr(), conv.weight.data_ptr(), z0.data_ptr(), dc, B, F, T, c0, s0, 0, 0, 0,
                   stt[0].data_ptr(), stt[1].data_ptr(), st)
        sc0, sv0 = self._finalize(bn, stt, B * Fi * Ti, dev)
        a = torch.empty_like(z0)
        L.bn_apply(z0.data_ptr(), sc0[0].data_ptr(), sc0[1].data_ptr(), HS, 0, a.data_ptr(), dc, B * Fi * Ti, c0, st)
        S["stem"] = dict(z=z0, sc=sc0, sv=sv0, Fo=Fi, To=Ti)
        for blk in self.blocks:
            a, Fi, Ti, R = self._block_train_fwd(blk, a, B, Fi, Ti)
            S["blocks"].append(R)
        logits, feat = self._head_train_fwd(a, B, Fi, Ti, S, dropout_mask)
        return logits, feat, S

    def _block_train_fwd(self, blk, a, B, Fi, Ti):
        return self._ir_block_train_fwd(blk, a, B, Fi, Ti)

    def _ir_block_train_fwd(self, blk, a, B, Fi, Ti):
        L = lib()
        dev, st = a.device, _stream()
        td, dc = self.tdtype, self.dcode
        R = {"inp": a, "Fi": Fi, "Ti": Ti}
        inp = a
        M = B * Fi * Ti
        if blk.expand is not None:
            z1 = torch.empty(B, Fi, Ti, blk.cexp, device=dev, dtype=td)
            stt = self._new_stats(blk.cexp, dev)
            self._gemm(inp, blk.expand[0].weight, z1, M, blk.cexp, blk.cin, stats=stt)
            sc1, sv1 = self._finalize(blk.expand[1], stt, M, dev)
            R.update(z1=z1, sc1=sc1, sv1=sv1)
            dw_in, dw_sc = z1, sc1
        else:
            dw_in, dw_sc = inp, None
        Fo, To = _conv_out(Fi, blk.k, blk.stride), _conv_out(Ti, blk.k, blk.stride)
        Mo = B * Fo * To
        z2 = torch.empty(B, Fo, To, blk.cexp, device=dev, dtype=td)
        stt = self._new_stats(blk.cexp, dev)
        wt = self._dw_weights(blk.dw[0], dev)
        L.dw_conv_fwd(dw_in.data_ptr(), wt.data_ptr(), z2.data_ptr(), dc, B, Fi, Ti, blk.cexp, blk.k, blk.stride,
                      _ptr(dw_sc[0]) if dw_sc is not None else 0, _ptr(dw_sc[1]) if dw_sc is not None else 0,
                      blk.act if dw_sc is not None else 0, 0, 0, 0, 0, stt[0].data_ptr(), stt[1].data_ptr(), st)
        sc2, sv2 = self._finalize(blk.dw[1], stt, Mo, dev)
        R.update(z2=z2, sc2=sc2, sv2=sv2, wt=wt, Fo=Fo, To=To)
        gate = None
        if blk.se is not None:
            Sq = blk.se.fc1.out_features
            pool = torch.zeros(B, blk.cexp, device=dev, dtype=torch.float32)
            L.bn_act_pool(z2.data_ptr(), sc2[0].data_ptr(), sc2[1].data_ptr(), blk.act, pool.data_ptr(),
                          1.0 / (Fo * To), dc, B, Fo * To, blk.cexp, st)
            gate, hidden = self._se_gate(blk.se, pool, 1.0, B, blk.cexp, dev)
            R.update(mean=pool, gate=gate, hidden=hidden)
        z3 = torch.empty(B, Fo, To, blk.cout, device=dev, dtype=td)
        stt = self._new_stats(blk.cout, dev)
        self._gemm(z2, blk.proj[0].weight, z3, Mo, blk.cout, blk.cexp, in_sc=sc2, in_act=blk.act, gate=gate,
                   rows_per_sample=Fo * To, stats=stt)
        sc3, sv3 = self._finalize(blk.proj[1], stt, Mo, dev)
        a = torch.empty(B, Fo, To, blk.cout, device=dev, dtype=td)
        L.bn_apply(z3.data_ptr(), sc3[0].data_ptr(), sc3[1].data_ptr(), 0, _ptr(inp) if blk.res else 0,
                   a.data_ptr(), dc, Mo, blk.cout, st)
        R.update(z3=z3, sc3=sc3, sv3=sv3)
        return a, Fo, To, R

    def _head_train_fwd(self, a, B, Fi, Ti, S, dropout_mask):
        L = lib()
        dev, st = a.device, _stream()
        td, dc = self.tdtype, self.dcode
        HS = ACT["hswish"]
        conv, bn = self.last[0], self.last[1]
        cl = conv.out_channels
        M = B * Fi * Ti
        zl = torch.empty(B, Fi, Ti, cl, device=dev, dtype=td)
        stt = self._new_stats(cl, dev)
        self._gemm(a, conv.weight, zl, M, cl, conv.in_channels, stats=stt)
        scl, svl = self._finalize(bn, stt, M, dev)
        feat = torch.zeros(B, cl, device=dev, dtype=torch.float32)
        L.bn_act_pool(zl.data_ptr(), scl[0].data_ptr(), scl[1].data_ptr(), HS, feat.data_ptr(), 1.0 / (Fi * Ti), dc, B,
                      Fi * Ti, cl, st)
        S["last"] = dict(inp=a, z=zl, sc=scl, sv=svl, Fi=Fi, Ti=Ti)
        # classifier: Linear -> Hardswish -> Dropout -> Linear.  The Hardswish and the dropout mask are applied on
        # the operand load of the second GEMM (in_act + per-row gate), so only the pre-activation is stored.
        n1 = self.fc1.out_features
        h_pre = torch.empty(B, n1, device=dev, dtype=torch.float32)
        self._gemm(feat, self.fc1.weight, h_pre, B, n1, cl, bias=self.fc1.bias, a_code=0, c_code=0)
        p = self.dropout_p
        if dropout_mask is None and p > 0:
            dropout_mask = torch.empty(B, n1, device=dev, dtype=torch.float32).bernoulli_(1.0 - p).div_(1.0 - p)
        ident = self._ident(n1, dev)
        logits = torch.empty(B, self.fc2.out_features, device=dev, dtype=torch.float32)
        self._gemm(h_pre, self.fc2.weight, logits, B, self.fc2.out_features, n1, in_sc=ident, in_act=HS,
                   gate=dropout_mask, rows_per_sample=1, bias=self.fc2.bias, a_code=0, c_code=0)
        S["head"] = dict(feat=feat, h_pre=h_pre, mask=dropout_mask, ident=ident)
        return logits, feat

    def _ident(self, n, dev):
        t = torch.empty(2, n, device=dev, dtype=torch.float32)
        t[0].fill_(1.0)
        t[1].zero_()
        return t

    # ------------------------------------------------------------------ backward
    def _block_modules(self):
        return list(self.model.features)[1:-1]

    def param_list(self):
        return [p for p in self.model.parameters()]

    def _wgrad(self, g, a, dW, db, M, N, K, in_sc=None, in_act=0, gate=None, rows_per_sample=1, g_code=None,
               a_code=None):
        g_code = self.dcode if g_code is None else g_code
        a_code = self.dcode if a_code is None else a_code
        args = (g.data_ptr(), g_code, a.data_ptr(), a_code, dW.data_ptr(), _ptr(db), M, N, K,
                _ptr(in_sc[0]) if in_sc is not None else 0, _ptr(in_sc[1]) if in_sc is not None else 0,
                in_act, _ptr(gate), rows_per_sample, _stream())
        use_tc = (self.gemm_impl != "simt" and db is None and g_code == a_code and M >= self.tc_min_rows
                  and K % 8 == 0 and N % 8 == 0)
        if use_tc:
            lib().pw_tc_wgrad(*args)
        else:
            lib().gemm_simt_wgrad(*args)

    def _bn_bwd(self, gA, gate, dpool, z, sc, sv, act, B, P, C, dgamma, dbeta, dev, code=None, sums=None):
        """two-pass BatchNorm(+activation) backward -> dz (same dtype/shape as z).  `sums`: the (s1, s2) accumulators when
        the reduce pass already happened elsewhere (SE blocks: eat_se_bn_bwd_reduce + eat_se_bn_bwd_combine)."""
        L = lib()
        st = _stream()
        code = self.dcode if code is None else code
        if sums is None:
            s = self._zero_pool.take(2, C, dev)
            L.bn_bwd_reduce(_ptr(gA), _ptr(gate), _ptr(dpool), z.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(),
                            sv[0].data_ptr(), sv[1].data_ptr(), act, code, B, P, C, s[0].data_ptr(), s[1].data_ptr(), st)
        else:
            s = sums
        coef = torch.empty(2, C, device=dev, dtype=torch.float32)
        L.bn_bwd_finalize(s[0].data_ptr(), s[1].data_ptr(), float(B * P), _ptr(dgamma), _ptr(dbeta),
                          coef[0].data_ptr(), coef[1].data_ptr(), C, st)
        dz = torch.empty_like(z)
        L.bn_bwd_apply(_ptr(gA), _ptr(gate), _ptr(dpool), z.data_ptr(), sc[0].data_ptr(), sc[1].data_ptr(),
                       sv[0].data_ptr(), sv[1].data_ptr(), act, coef[0].data_ptr(), coef[1].data_ptr(), dz.data_ptr(),
                       code, B, P, C, st)
        return dz

    def _block_bwd(self, blk, R, dy, G, B):
        return self._ir_block_bwd(blk, R, dy, G, B)

    def _ir_block_bwd(self, blk, R, dy, G, B):
        """backward of one InvertedResidual; returns the gradient w.r.t. the block input"""
        L = lib()
        st = _stream()
        dev = dy.device
        td, dc = self.tdtype, self.dcode
        Fi, Ti, Fo, To = R["Fi"], R["Ti"], R["Fo"], R["To"]
        Pi, Po = Fi * Ti, Fo * To
        gate = R.get("gate")
        # project: BN3 (no activation)
        dz3 = self._bn_bwd(dy, None, None, R["z3"], R["sc3"], R["sv3"], 0, B, Po, blk.cout,
                           G[blk.proj[1].weight], G[blk.proj[1].bias], dev)
        fork = self._fork if self._fork is not None else _Fork(dev, False)
        fork.run(lambda: self._wgrad(dz3, R["z2"], G[blk.proj[0].weight], None, B * Po, blk.cout, blk.cexp, in_sc=R["sc2"],
                                     in_act=blk.act, gate=gate, rows_per_sample=Po), dz3)
        dp = torch.empty_like(R["z2"])
        self._gemm(dz3, blk.proj[0].weight, dp, B * Po, blk.cexp, blk.cout, w_trans=True)
        dpool = None
        if blk.se is not None:
            Sq = blk.se.fc1.out_features
            dgate = torch.zeros(B, blk.cexp, device=dev, dtype=torch.float32)
            if self.se_fused:
                # slices of a sample's pixels = CTAs per sample: ~EAT_SE_PARTS_PER_SM CTAs per SM over the batch, at least
                # ~16 pixels per slice.  Every slice writes 4 x C partial sums, so few, long slices are preferred.
                parts = max(1, min(32, (148 * self.se_parts_per_sm) // B, (Po + 15) // 16))
                part = torch.empty(parts, 4, B, blk.cexp, device=dev, dtype=torch.float32)
                L.se_bn_bwd_reduce(dp.data_ptr(), R["z2"].data_ptr(), R["sc2"][0].data_ptr(), R["sc2"][1].data_ptr(),
                                   R["sv2"][0].data_ptr(), blk.act, dgate.data_ptr(), part.data_ptr(), parts, dc, B, Po,
                                   blk.cexp, st)
            else:
                L.se_bwd_reduce(dp.data_ptr(), R["z2"].data_ptr(), R["sc2"][0].data_ptr(), R["sc2"][1].data_ptr(),
                                blk.act, dgate.data_ptr(), dc, B, Po, blk.cexp, st)
            du2 = torch.empty(B, blk.cexp, device=dev, dtype=torch.float32)
            du1 = torch.empty(B, Sq, device=dev, dtype=torch.float32)
            dpool = torch.empty(B, blk.cexp, device=dev, dtype=torch.float32)
            L.se_fc_bwd(dgate.data_ptr(), gate.data_ptr(), R["hidden"].data_ptr(), blk.se.fc1.weight.data_ptr(),
                        blk.se.fc2.weight.data_ptr(), 1.0 / Po, du2.data_ptr(), du1.data_ptr(), dpool.data_ptr(),
                        B, blk.cexp, Sq, st)
            fork.run(lambda: (self._wgrad(du2, R["hidden"], G[blk.se.fc2.weight], G[blk.se.fc2.bias], B, blk.cexp, Sq, g_code=0, a_code=0),
                              self._wgrad(du1, R["mean"], G[blk.se.fc1.weight], G[blk.se.fc1.bias], B, Sq, blk.cexp, g_code=0, a_code=0)),
                     du2, du1)
        # depthwise: BN2 + activation (+ SE gate / squeeze gradient composed on the fly)
        sums2 = None
        if blk.se is not None and self.se_fused:
            sums2 = self._zero_pool.take(2, blk.cexp, dev)
            L.se_bn_bwd_combine(part.data_ptr(), parts, gate.data_ptr(), dpool.data_ptr(), R["sv2"][1].data_ptr(), B,
                                blk.cexp, sums2[0].data_ptr(), sums2[1].data_ptr(), st)
        dz2 = self._bn_bwd(dp, gate, dpool, R["z2"], R["sc2"], R["sv2"], blk.act, B, Po, blk.cexp,
                           G[blk.dw[1].weight], G[blk.dw[1].bias], dev, sums=sums2)
        has_exp = blk.expand is not None
        dw_in = R["z1"] if has_exp else R["inp"]
        sc1 = R["sc1"] if has_exp else None
        fork.run(lambda: L.dw_conv_wgrad(dz2.data_ptr(), dw_in.data_ptr(), _ptr(sc1[0]) if has_exp else 0,
                                         _ptr(sc1[1]) if has_exp else 0, blk.act if has_exp else 0,
                                         G[blk.dw[0].weight].data_ptr(), 0, dc, B, Fi, Ti, blk.cexp, blk.k, blk.stride,
                                         _stream()), dz2)
        da1 = torch.empty_like(dw_in)
        sums1 = None
        if has_exp and self.dgrad_bnred and blk.stride == 2 and dc == 0 and blk.k in (3, 5):
            # the expand BatchNorm's reduce pass rides in the epilogue of the kernel that produces its upstream gradient
            sums1 = self._zero_pool.take(2, blk.cexp, dev)
            L.dw_conv_dgrad_bnred(dz2.data_ptr(), R["wt"].data_ptr(), 0, da1.data_ptr(), R["z1"].data_ptr(),
                                  R["sc1"][0].data_ptr(), R["sc1"][1].data_ptr(), R["sv1"][0].data_ptr(),
                                  R["sv1"][1].data_ptr(), blk.act, sums1[0].data_ptr(), sums1[1].data_ptr(), dc, B, Fi, Ti,
                                  blk.cexp, blk.k, blk.stride, st)
        else:
            # without an expand stage the depthwise input IS the block input: fold the residual gradient in
            L.dw_conv_dgrad(dz2.data_ptr(), R["wt"].data_ptr(), 0, _ptr(dy) if (blk.res and not has_exp) else 0,
                            da1.data_ptr(), dc, B, Fi, Ti, blk.cexp, blk.k, blk.stride, st)
        if has_exp:
            dz1 = self._bn_bwd(da1, None, None, R["z1"], R["sc1"], R["sv1"], blk.act, B, Pi, blk.cexp,
                               G[blk.expand[1].weight], G[blk.expand[1].bias], dev, sums=sums1)
            fork.run(lambda: self._wgrad(dz1, R["inp"], G[blk.expand[0].weight], None, B * Pi, blk.cexp, blk.cin), dz1)
            dinp = torch.empty_like(R["inp"])
            self._gemm(dz1, blk.expand[0].weight, dinp, B * Pi, blk.cin, blk.cexp, w_trans=True,
                       res=dy if blk.res else None)
            dy = dinp
        else:
            dy = da1
        return dy

    def _backward(self, S, dlogits, on_ready=None):
        """-> dict {parameter: fp32 gradient view into one flat arena} (arena returned under key None).
        on_ready(flat, i): called after each stage with the index i of the first parameter (model.parameters() order)
        whose gradient is final -- everything from i to the end is -- so a data-parallel trainer can start reducing."""
        L = lib()
        st = _stream()
        dev = dlogits.device
        td, dc = self.tdtype, self.dcode
        B = S["B"]
        HS = ACT["hswish"]
        self._zero_pool = _ZeroPool()
        params = self.param_list()
        flat = torch.zeros(sum(p.numel() for p in params), device=dev, dtype=torch.float32)
        G, off = {}, 0
        for p in params:
            G[p] = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        dlogits = dlogits.float().contiguous()
        fork = self._fork = _Fork(dev, self.fork_wgrad)
        pidx = {id(p): i for i, p in enumerate(params)}

        def done(module):
            fork.join()                                  # this stage's weight gradients are complete
            if on_ready is not None:
                on_ready(flat, min(pidx[id(p)] for p in module.parameters()))

        # ---- classifier
        H = S["head"]
        n1, ncls, cl = self.fc1.out_features, self.fc2.out_features, self.fc1.in_features
        fork.run(lambda: self._wgrad(dlogits, H["h_pre"], G[self.fc2.weight], G[self.fc2.bias], B, ncls, n1, in_sc=H["ident"],
                                     in_act=HS, gate=H["mask"], rows_per_sample=1, g_code=0, a_code=0), dlogits)
        dh = torch.empty(B, n1, device=dev, dtype=torch.float32)
        self._gemm(dlogits, self.fc2.weight, dh, B, n1, ncls, a_code=0, c_code=0, w_trans=True)
        dpre = torch.empty_like(dh)
        L.act_bwd(dh.data_ptr(), H["h_pre"].data_ptr(), _ptr(H["mask"]), HS, dpre.data_ptr(), dh.numel(), st)
        fork.run(lambda: self._wgrad(dpre, H["feat"], G[self.fc1.weight], G[self.fc1.bias], B, n1, cl, g_code=0, a_code=0), dpre)
        dfeat = torch.empty(B, cl, device=dev, dtype=torch.float32)
        self._gemm(dpre, self.fc1.weight, dfeat, B, cl, n1, a_code=0, c_code=0, w_trans=True)
        done(self.fc1)

        # ---- last 1x1 conv (+BN+Hardswish, global average pool)
        Ls = S["last"]
        P = Ls["Fi"] * Ls["Ti"]
        conv, bn = self.last[0], self.last[1]
        dpool = dfeat.mul_(1.0 / P)      # gradient of the spatial mean, broadcast inside the BN-backward kernels
        dz = self._bn_bwd(None, None, dpool, Ls["z"], Ls["sc"], Ls["sv"], HS, B, P, cl, G[bn.weight], G[bn.bias], dev)
        fork.run(lambda: self._wgrad(dz, Ls["inp"], G[conv.weight], None, B * P, cl, conv.in_channels), dz)
        dy = torch.empty_like(Ls["inp"])
        self._gemm(dz, conv.weight, dy, B * P, conv.in_channels, cl, w_trans=True)
        done(self.last)

        # ---- blocks, last to first
        for blk, R, mod in zip(reversed(self.blocks), reversed(S["blocks"]), reversed(self._block_modules())):
            dy = self._block_bwd(blk, R, dy, G, B)
            done(mod)
        # ---- stem
        St = S["stem"]
        conv, bn = self.stem[0], self.stem[1]
        c0 = conv.out_channels
        dz0 = self._bn_bwd(dy, None, None, St["z"], St["sc"], St["sv"], HS, B, St["Fo"] * St["To"], c0, G[bn.weight],
                           G[bn.bias], dev)
        L.stem_wgrad(dz0.data_ptr(), dc, S["x"].data_ptr(), G[conv.weight].data_ptr(), B, S["F"], S["T"], c0,
                     conv.stride[0], st)
        fork.join()
        self._fork = None
        G[None] = flat
        return G
