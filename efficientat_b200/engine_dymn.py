"""DyMN execution engine (eval forward): ContextGen -> DynamicConv 1x1 (tcgen05, kernel mix fused into the weight
staging) -> BN+act -> DynamicConv depthwise (per-sample tap tables) + BN + DyReLU-B + CoordAtt in one kernel ->
DynamicConv 1x1 + BN (+ residual).  Reference models/dymn/dy_block.py:390-409, models/dymn/model.py:157-200.

The training step (batch-statistics forward + backward) of DyMN is not implemented yet: calling the model in
training mode raises NotImplementedError (never a silent PyTorch fallback)."""
import torch

from ._lib import lib
from .engine import ACT, MNEngine, _Layer, _conv_out, _ptr, _stream


class DyMNEngine(MNEngine):
    def _plan(self):
        from .models.dymn.dy_block import DY_Block
        from .models.mn.block_types import ConvNormActivation, InvertedResidual
        m = self.model
        self.stem, self.last = m.in_c, m.out_c
        self.blocks = []
        for blk in m.layers:
            L = _Layer()
            L.dy = isinstance(blk, DY_Block)
            cnf = blk.cnf
            L.act = ACT["hswish"] if cnf.use_hs else ACT["relu"]
            L.k, L.stride = cnf.kernel, cnf.stride
            L.cin, L.cexp, L.cout = cnf.input_channels, cnf.expanded_channels, cnf.out_channels
            L.res = blk.use_res_connect
            if L.dy:
                L.m = blk
                L.has_exp = cnf.expanded_channels != cnf.input_channels
                L.H = blk.context_dim
            else:
                assert isinstance(blk, InvertedResidual)
                subs = [s for s in blk.block if isinstance(s, ConvNormActivation)]
                L.expand = subs[0] if len(subs) == 3 else None
                L.dw, L.proj, L.se = subs[-2], subs[-1], None
            self.blocks.append(L)
        self.fc1, self.fc2 = m.classifier[2], m.classifier[5]
        self.dropout_p = m.classifier[4].p

    def forward(self, x, return_fmaps=False):
        if not x.is_cuda:
            raise RuntimeError("efficientat_b200 models run on CUDA (sm_100a) only; got a CPU tensor")
        if x.dim() != 4 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape [B, 1, F, T], got {tuple(x.shape)}")
        if self.model.training:
            raise NotImplementedError("DyMN: the fused training step (batch-statistics forward + backward) is not "
                                      "implemented yet; use model.eval() (inference) -- see DESIGN.md section 7")
        logits, feat, fmaps = self._forward_eval(x.detach(), return_fmaps)
        return logits, feat, fmaps

    # ------------------------------------------------------------------
    def _dy_block_eval(self, blk, a, B, Fi, Ti):
        L = lib()
        st = _stream()
        dev = a.device
        td, dc = self.tdtype, self.dcode
        m = blk.m
        H = blk.H
        cg = m.context_gen
        P = Fi + Ti
        f32 = torch.float32
        # ---- ContextGen (dy_block.py:235-254)
        g = torch.empty(B, P, blk.cin, device=dev, dtype=f32)
        L.ctx_pool(a.data_ptr(), dc, g.data_ptr(), B, Fi, Ti, blk.cin, st)
        hcat = torch.empty(B * P, H, device=dev, dtype=f32)
        self._gemm(g, cg.joint_conv.weight, hcat, B * P, H, blk.cin, sc=self._fold(cg.joint_norm, dev), act=ACT["hswish"],
                   a_code=0, c_code=0)
        h_c = torch.zeros(B, H, device=dev, dtype=f32)
        L.bn_act_pool(hcat.data_ptr(), 0, 0, 0, h_c.data_ptr(), 1.0 / P, 0, B, P, H, st)
        s = blk.stride
        Fo, To = _conv_out(Fi, blk.k, s), _conv_out(Ti, blk.k, s)
        hf = torch.empty(B, Fo, H, device=dev, dtype=f32)
        ht = torch.empty(B, To, H, device=dev, dtype=f32)
        L.seq_pool(hcat.data_ptr(), hf.data_ptr(), B, P, 0, Fi, H, s, st)
        L.seq_pool(hcat.data_ptr(), ht.data_ptr(), B, P, Fi, Ti, H, s, st)
        ca_f = torch.empty(B, Fo, blk.cexp, device=dev, dtype=f32)
        ca_t = torch.empty(B, To, blk.cexp, device=dev, dtype=f32)
        SIG = 3
        self._gemm(hf, cg.conv_f.weight, ca_f, B * Fo, blk.cexp, H, bias=cg.conv_f.bias, act=SIG, a_code=0, c_code=0)
        self._gemm(ht, cg.conv_t.weight, ca_t, B * To, blk.cexp, H, bias=cg.conv_t.bias, act=SIG, a_code=0, c_code=0)

        def attention(dc_mod):
            att = torch.empty(B, dc_mod.k, device=dev, dtype=f32)
            lin = dc_mod.residuals[0]
            L.dyconv_att(h_c.data_ptr(), lin.weight.data_ptr(), lin.bias.data_ptr(), float(dc_mod.temperature),
                         att.data_ptr(), B, H, dc_mod.k, st)
            return att

        # ---- expand: DynamicConv 1x1 + BN + act
        inp = a
        if blk.has_exp:
            att = attention(m.exp_conv)
            sc = self._fold(m.exp_norm, dev)
            e = torch.empty(B, Fi, Ti, blk.cexp, device=dev, dtype=td)
            L.pw_tc_dyn_fwd(inp.data_ptr(), dc, m.exp_conv.weight.data_ptr(), att.data_ptr(), m.exp_conv.k, e.data_ptr(),
                            B * Fi * Ti, blk.cexp, blk.cin, Fi * Ti, 0, 0, 0, sc[0].data_ptr(), sc[1].data_ptr(), blk.act,
                            0, 0, 0, st)
        else:
            e = inp
        # ---- depthwise DynamicConv + BN + DyReLU-B + CoordAtt
        att = attention(m.depth_conv)
        kk = blk.k * blk.k
        wt = torch.empty(B, kk, blk.cexp, device=dev, dtype=f32)
        L.dyconv_mix_dw(m.depth_conv.weight.data_ptr(), att.data_ptr(), wt.data_ptr(), B, blk.cexp, blk.k,
                        m.depth_conv.k, st)
        coef = m.depth_act.coef_net[0]
        theta = torch.empty(B, 4 * blk.cexp, device=dev, dtype=f32)
        self._gemm(h_c, coef.weight, theta, B, 4 * blk.cexp, H, bias=coef.bias, act=SIG, a_code=0, c_code=0)
        sc = self._fold(m.depth_norm, dev)
        d = torch.empty(B, Fo, To, blk.cexp, device=dev, dtype=td)
        L.dw_conv_fwd_dy(e.data_ptr(), wt.data_ptr(), kk * blk.cexp, d.data_ptr(), dc, B, Fi, Ti, blk.cexp, blk.k, s,
                         0, 0, 0, sc[0].data_ptr(), sc[1].data_ptr(), theta.data_ptr(), m.depth_act.lambdas.data_ptr(),
                         m.depth_act.init_v.data_ptr(), ca_f.data_ptr(), ca_t.data_ptr(), 0, 0, st)
        # ---- project: DynamicConv 1x1 + BN (+ residual)
        att = attention(m.proj_conv)
        sc = self._fold(m.proj_norm, dev)
        o = torch.empty(B, Fo, To, blk.cout, device=dev, dtype=td)
        L.pw_tc_dyn_fwd(d.data_ptr(), dc, m.proj_conv.weight.data_ptr(), att.data_ptr(), m.proj_conv.k, o.data_ptr(),
                        B * Fo * To, blk.cout, blk.cexp, Fo * To, 0, 0, 0, sc[0].data_ptr(), sc[1].data_ptr(), 0,
                        _ptr(inp) if blk.res else 0, 0, 0, st)
        return o, Fo, To

    def _forward_eval(self, x, return_fmaps=False):
        L = lib()
        dev = x.device
        st = _stream()
        td, dc = self.tdtype, self.dcode
        x = x.float().contiguous()
        B, _, F, T = x.shape
        fmaps = [] if return_fmaps else None

        def keep(t, f, tt, c):
            if fmaps is not None:
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
            if blk.dy:
                a, Fi, Ti = self._dy_block_eval(blk, a, B, Fi, Ti)
            else:
                a, Fi, Ti = self._ir_block_eval(blk, a, B, Fi, Ti)
            keep(a, Fi, Ti, blk.cout)
        logits, feat, z = self._head_eval(a, B, Fi, Ti)
        keep(z, Fi, Ti, self.last[0].out_channels)
        return logits, feat, fmaps
