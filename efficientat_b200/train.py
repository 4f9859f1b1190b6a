"""Device-resident AudioSet training step: the body of the reference's loop, ex_audioset.py:135-199
(mel -> mixup -> model -> hard-label + distillation BCE -> backward -> Adam), as a chain of this
package's kernels with one optional NCCL all-reduce of the flat gradient arena (the data-parallel
semantics of ex_pl_audioset.py:287-293: replicas, per-replica BatchNorm, gradient mean).

No host synchronisation happens inside `step`; the loss comes back as a device tensor."""
import torch

from ._lib import lib
from .helpers.utils import mixup as draw_mixup


def _stream():
    return torch.cuda.current_stream().cuda_stream


class AudioSetTrainer:
    def __init__(self, model, mel, lr=8e-4, kd_lambda=0.1, mixup_alpha=0.3, weight_decay=0.0, adamw=False,
                 betas=(0.9, 0.999), eps=1e-8, process_group=None):
        self.model, self.mel = model, mel
        self.engine = model.engine()
        self.lr, self.kd_lambda, self.mixup_alpha = lr, kd_lambda, mixup_alpha
        self.weight_decay, self.adamw, self.betas, self.eps = weight_decay, adamw, betas, eps
        self.pg = process_group
        self.world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self._flatten()
        self.steps = 0

    def _flatten(self):
        """Re-point every parameter into one contiguous fp32 arena (order = model.parameters()), so the
        optimiser is one kernel and the gradient all-reduce one collective.  state_dict()/load_state_dict()
        keep working (they copy in place)."""
        params = self.engine.param_list()
        dev = params[0].device
        n = sum(p.numel() for p in params)
        flat = torch.empty(n, device=dev, dtype=torch.float32)
        off = 0
        for p in params:
            k = p.numel()
            flat[off:off + k].copy_(p.data.reshape(-1))
            p.data = flat[off:off + k].view(p.shape)
            off += k
        self.flat_p = flat
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        if self.world > 1:        # replicas start from rank 0's parameters (DDP constructor semantics)
            torch.distributed.broadcast(self.flat_p, 0, group=self.pg)

    def forward_backward(self, wave, y, teacher=None, perm=None, lam=None):
        """-> (loss_acc fp64[2] device = weighted label / distillation losses, flat gradient arena)."""
        L = lib()
        st = _stream()
        B = wave.shape[0]
        spec = self.mel(wave.reshape(B, -1))                       # [B, n_mels, T]
        if self.mixup_alpha and perm is None:
            perm, lam = draw_mixup(B, self.mixup_alpha)
        if perm is not None:
            perm_d = perm.to(device=spec.device, dtype=torch.int32, non_blocking=True)
            lam_d = lam.to(device=spec.device, dtype=torch.float32, non_blocking=True)
            mixed = torch.empty_like(spec)
            L.mixup(spec.data_ptr(), perm_d.data_ptr(), lam_d.data_ptr(), mixed.data_ptr(), B,
                    spec.shape[1] * spec.shape[2], st)
            spec = mixed
        else:
            perm_d = lam_d = None
        logits, _, saved = self.engine._forward_train(spec.unsqueeze(1))
        dlogits = torch.empty_like(logits)
        loss_acc = torch.zeros(2, device=spec.device, dtype=torch.float64)
        L.bce_kd_loss(logits.data_ptr(), y.data_ptr(), teacher.data_ptr() if teacher is not None else 0,
                      perm_d.data_ptr() if perm_d is not None else 0, lam_d.data_ptr() if lam_d is not None else 0,
                      self.kd_lambda, B, logits.shape[1], dlogits.data_ptr(), loss_acc.data_ptr(), st)
        grads = self.engine._backward(saved, dlogits)
        return loss_acc, grads[None]

    def step(self, wave, y, teacher=None, perm=None, lam=None):
        self.model.train()
        self.mel.train()
        loss_acc, flat_g = self.forward_backward(wave, y, teacher, perm, lam)
        if self.world > 1:
            torch.distributed.all_reduce(flat_g, group=self.pg)        # one collective per step (sum)
        self.steps += 1
        lib().adam_step(self.flat_p.data_ptr(), flat_g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                        flat_g.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                        1 if self.adamw else 0, self.steps, 1.0 / self.world, _stream())
        return loss_acc
