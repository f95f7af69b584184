#!/usr/bin/env python
"""Which kernel class adds the bf16 gradient noise?  Runs the tiny smoke model on the GPU with ONE class of entry points at a
time replaced by its fp64-math torch reference (tests/ref_ops.py: same bf16 storage, ideal arithmetic) and prints the worst
temporal-bias-table error and the worst tensor overall against the fp64 oracle.  TEST / ANALYSIS CODE, never a product path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stg-cma_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402
import model  # noqa: E402
import ref_ops  # noqa: E402
import stgcma.lib as lib  # noqa: E402
from oracle import swin_ave_oracle as O  # noqa: E402

kw = dict(img_size=112, num_frames=3, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
          adapter_mlp_ratio=[0.25, 0.25, 0.125], label_dim=7)
CLASSES = {
    "none": [],
    "gemm": ["gemm"], "gemm_tn": ["gemm_tn"], "mlp": ["mlp_fwd", "mlp_bwd"], "layernorm": ["layernorm_fwd", "layernorm_bwd"],
    "window": ["window_attn_fwd", "window_attn_bwd"], "temporal": ["temporal_attn_fwd", "temporal_attn_bwd"],
    "xwin": ["xmodal_window_fwd", "xmodal_window_bwd"], "xglobal": ["xmodal_global_fwd", "xmodal_global_bwd"],
    "patch_embed": ["patch_embed_fwd"],
}
CLASSES["all"] = sorted({n for v in CLASSES.values() for n in v})


def run(seed, names):
    cfg = O.SwinAVEConfig(**kw)
    sd = O.make_state(cfg, seed=seed, bf16_exact=True)
    a, v, labels = O.make_inputs(cfg, 2, seed=seed)
    ref_out, ref_loss, ref_g = O.train_step({k: (t.double() if t.is_floating_point() else t) for k, t in sd.items()},
                                            a.double(), v.double(), labels.double(), cfg, "fusion")
    saved = {n: getattr(lib, n) for n in names}
    for n in names:
        setattr(lib, n, getattr(ref_ops, n))
    try:
        m = model.Swin_AVE.SwinTransformer2D_Adapter_New(ftmode="fusion", drop_path_rate=0.0, compute_dtype=torch.bfloat16,
                                                         patch_size=[1, 4, 4], **kw).cuda()
        m.load_state_dict(sd, strict=True)
        m.train()
        m.mlp_head[1].eval()
        for n, p in m.named_parameters():
            p.requires_grad_(O.is_trainable(n))
        out = m(a.cuda(), v.cuda(), "fusion")
        torch.nn.CrossEntropyLoss()(out, labels.cuda()).backward()
        torch.cuda.synchronize()
    finally:
        for n, f in saved.items():
            setattr(lib, n, f)
    worst_t, worst = (0, 0, ""), (0, "")
    for n, p in m.named_parameters():
        if p.requires_grad and "gate_" not in n:
            g, r = p.grad.cpu().double(), ref_g[n]
            e2 = float((g - r).norm() / r.norm().clamp_min(1e-30))
            em = float((g - r).abs().max() / r.abs().max().clamp_min(1e-30))
            if "temporal_position_bias_table" in n and em > worst_t[1]:
                worst_t = (e2, em, n)
            if e2 > worst[0]:
                worst = (e2, n)
    return worst_t, worst


for seed in (3, 11):
    for cname, names in CLASSES.items():
        try:
            wt, w = run(seed, names)
            print(f"seed {seed:2d} ideal[{cname:11s}]: temporal table L2 {wt[0]:.3e} max-abs {wt[1]:.3e} ({wt[2][:40]})   worst L2 {w[0]:.3e} {w[1][:50]}", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"seed {seed} ideal[{cname}]: FAILED {type(e).__name__}: {str(e)[:200]}", flush=True)
