import contextlib, io, torch, sys
sys.path.insert(0,'.')
from efficientat_b200.models.mn.model import get_model
from efficientat_b200.models.preprocess import AugmentMelSTFT
from efficientat_b200.synth import synth_labels, synth_state_, synth_waveform
from oracle import mel_oracle, net_oracle
import os
dev=torch.device('cuda',0)
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    model = synth_state_(get_model(width_mult=1.0, verbose=False), seed=3)
    mel = AugmentMelSTFT(freqm=0, timem=0)
sd = {k: v.clone() for k, v in model.state_dict().items()}
model.to(dev).train(); model.classifier[4].p = 0.0; model.engine().dropout_p = 0.0
if os.environ.get('GEMM'): model.engine().gemm_impl = os.environ['GEMM']
mel.to(dev).eval()
wave = synth_waveform(2, 32000, seed=5); y = synth_labels(2, 527, seed=6)
spec = mel(wave.to(dev)); logits,_ = model(spec.unsqueeze(1))
loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, y.to(dev)); loss.backward()
names=[k for k,_ in model.named_parameters()]
for k in names: sd[k].requires_grad_(True)
spec_ref = mel_oracle.mel_forward(wave)
ref_logits,_ = net_oracle.mn_forward(sd, spec_ref.unsqueeze(1), training=True)
ref_loss = torch.nn.functional.binary_cross_entropy_with_logits(ref_logits, y); ref_loss.backward()
gp=dict(model.named_parameters())
errs=[]
for k in names:
    a,b=gp[k].grad.cpu(), sd[k].grad
    errs.append(((a-b).norm()/(b.norm()+1e-12)).item())
import numpy as np
order=np.argsort(errs)[::-1]
print('logit err', (logits.detach().cpu()-ref_logits.detach()).abs().max().item())
for i in order[:8]: print(names[i], errs[i], gp[names[i]].grad.norm().item())
print('median', float(np.median(errs)), 'stem', errs[names.index('features.0.0.weight')])
