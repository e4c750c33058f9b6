"""Debug helper: per-parameter gradient error of the CUDA path vs the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import micformer_oracle as O
from micformer_b200.models.MICFormer_self import Head, MicFormer
from micformer_b200.loss.dice import MDiceLoss

cfgname, S = sys.argv[1], int(sys.argv[2])
cfg = getattr(O, cfgname)
sd = O.synth_state_dict(cfg, seed=7)
x, lab = O.synth_inputs(1, S, cfg.num_classes, seed=9)
head = Head(embed_dim=cfg.embed_dim, num_classes=cfg.num_classes, window_size=cfg.window_size)
if tuple(cfg.depths) != (2, 2, 6, 2) or tuple(cfg.num_heads) != (3, 6, 12, 24):
    head.swin = MicFormer(window_size=cfg.window_size, in_chans=1, embed_dim=cfg.embed_dim, depths=list(cfg.depths), num_heads=list(cfg.num_heads))
head.load_state_dict(sd); head = head.cuda().eval()
y = head(x.cuda()); loss = MDiceLoss()(y, lab.cuda()); loss.backward()
logits, loss_ref, grads = O.train_step(x, lab, sd, cfg)
print("logits rel", float((y.detach().cpu()-logits).abs().max()/logits.abs().max()), "loss", float(loss), float(loss_ref))
errs = []
for k, p in head.named_parameters():
    if grads[k] is None: continue
    e = float((p.grad.cpu()-grads[k]).norm()/(grads[k].norm()+1e-30))
    errs.append((e, k, float(grads[k].norm())))
errs.sort(reverse=True)
for e, k, n in errs[:25]: print(f"{e:.3e}  |g|={n:.3e}  {k}")
print("median", errs[len(errs)//2][0])
