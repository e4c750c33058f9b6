import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import micformer_oracle as O
from micformer_b200.models.MICFormer_self import Head
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = O.TRAIN
for wseed, default_init in ((7, False), (0, True)):
    if default_init:
        torch.manual_seed(0); head = Head(embed_dim=48, num_classes=8); sd = {k: v.detach().clone() for k, v in head.state_dict().items()}
    else:
        sd = O.synth_state_dict(cfg, seed=wseed); head = Head(embed_dim=48, num_classes=8); head.load_state_dict(sd)
    head = head.cuda().eval()
    x, _ = O.synth_inputs(1, S, 8, seed=9)
    with torch.no_grad():
        ref = O.head_forward(x, sd, cfg)
        y = head(x.cuda()).cpu()
    print(f"S={S} default_init={default_init} exact_tail='{os.environ.get('MICFORMER_DEBUG_EXACT_TAIL','')}' logits rel err {float((y-ref).abs().max()/ref.abs().max()):.3e}  rms rel {float((y-ref).norm()/ref.norm()):.3e}", flush=True)
