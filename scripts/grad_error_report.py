#!/usr/bin/env python
"""Per-tensor bf16 gradient error of the tiny smoke model against the fp64 oracle (where does the 2e-2 budget go?)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stg-cma_b200")]
import torch  # noqa: E402
import model  # noqa: E402
from oracle import swin_ave_oracle as O  # noqa: E402

kw = dict(img_size=112, num_frames=3, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
          adapter_mlp_ratio=[0.25, 0.25, 0.125], label_dim=7)
cfg = O.SwinAVEConfig(**kw)
sd = O.make_state(cfg, seed=3, bf16_exact=True)
a, v, labels = O.make_inputs(cfg, 2, seed=3)
ref_out, ref_loss, ref_g = O.train_step({k: (t.double() if t.is_floating_point() else t) for k, t in sd.items()},
                                        a.double(), v.double(), labels.double(), cfg, "fusion")
m = model.Swin_AVE.SwinTransformer2D_Adapter_New(ftmode="fusion", drop_path_rate=0.0, compute_dtype=torch.bfloat16,
                                                 patch_size=[1, 4, 4], **kw).cuda()
m.load_state_dict(sd, strict=True)
m.train()
m.mlp_head[1].eval()
for n, p in m.named_parameters():
    p.requires_grad_(O.is_trainable(n))
out = m(a.cuda(), v.cuda(), "fusion")
torch.nn.CrossEntropyLoss()(out, labels.cuda()).backward()
rows = []
for n, p in m.named_parameters():
    if p.requires_grad and "gate_" not in n:
        g, r = p.grad.cpu().double(), ref_g[n]
        rows.append((float((g - r).norm() / r.norm().clamp_min(1e-30)), float((g - r).abs().max() / r.abs().max().clamp_min(1e-30)), n, tuple(g.shape)))
rows.sort(reverse=True)
for e2, em, n, sh in rows[:14]:
    print(f"{e2:.3e}  {em:.3e}  {n}  {sh}")
print("median rel-L2:", sorted(r[0] for r in rows)[len(rows) // 2])
