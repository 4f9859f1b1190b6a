"""torch.autograd glue for the training step: one Function for the whole network, so that the reference's
training loop (`loss.backward()`, ex_audioset.py:197) drives the hand-written backward chain in
efficientat_b200.engine unchanged.  Parameter gradients come back as views of one flat fp32 arena."""
import torch


class _MNTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, *params):
        logits, feat, saved = engine._forward_train(x)
        ctx.engine = engine
        ctx.saved = saved
        ctx.mark_non_differentiable(feat)
        return logits, feat

    @staticmethod
    def backward(ctx, dlogits, _dfeat):
        grads = ctx.engine._backward(ctx.saved, dlogits)
        ctx.saved = None
        out = [grads[p] if p.requires_grad else None for p in ctx.engine.param_list()]
        return (None, None) + tuple(out)


def mn_train_forward(engine, x, needs_grad):
    if not needs_grad:
        logits, feat, _ = engine._forward_train(x)
        return logits, feat
    if x.requires_grad:
        raise NotImplementedError("gradients w.r.t. the input spectrogram are not implemented")
    return _MNTrainFn.apply(engine, x, *engine.param_list())
