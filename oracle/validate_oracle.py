"""TEST INFRASTRUCTURE ONLY.  Pins oracle/swin_ave_oracle.py to the UNMODIFIED reference.

Runs only where /root/reference exists (the build container).  For each case it
  1. builds the reference `SwinTransformer2D_Adapter_New` (AVE/model/Swin_AVE.py:1129) through
     oracle/shim.py with drop_path_rate=0 and mlp_head[1] (Dropout) in eval,
  2. loads the oracle's seeded synthetic state with load_state_dict(strict=True) -- this
     also proves the names / shapes / buffers (incl. the integer index buffers and the
     0/-100 shift masks, compared bit-exactly) match,
  3. applies the caller's freeze rule (traintest_adapt_ave29.py:38-61), runs forward +
     CrossEntropyLoss(soft targets) + backward on both sides,
  4. reports max |diff| / max |ref| for logits and every trainable gradient.

    python -m oracle.validate_oracle            # prints a table, exits non-zero on mismatch
"""
import sys
import time

import torch

from . import shim
from . import swin_ave_oracle as O

CASES = [
    # name, cfg kwargs, batch, mode
    ("tiny3stage_fusion", dict(img_size=112, num_frames=3, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
                               adapter_mlp_ratio=[0.25, 0.25, 0.125], label_dim=7), 2, "fusion"),
    ("tiny3stage_multimodal", dict(img_size=112, num_frames=3, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
                                   adapter_mlp_ratio=[0.25, 0.25, 0.125], label_dim=7), 2, "multimodal"),
    ("swin_t_T2_fusion", dict(img_size=224, num_frames=2, label_dim=29, **O.PRESETS["swin_t"]), 1, "fusion"),
    # BASELINE.json configs[0]: AVE 29-class Swin-T, batch 2, 10 frames 224^2 (the reference's CPU-runnable case)
    ("swin_t_T10_B2_fusion", dict(img_size=224, num_frames=10, label_dim=29, **O.PRESETS["swin_t"]), 2, "fusion"),
    # the configuration the authors publish (run_adapt_ave29.py:156-165) at FULL depth [2, 2, 18, 2]: heads 4 / 8 / 16 / 32
    # (window attention in head groups of 2), r = 16 / 32 / 32 / 64, C = 128 .. 1024
    ("swin_b_T2_B1_fusion", dict(img_size=224, num_frames=2, label_dim=29, **O.PRESETS["swin_b"]), 1, "fusion"),
    # run_adapt_ave29.py:171-181 at full depth: C = 192 .. 1536, heads 6 .. 48, r = 96 at every stage (ratios .5 / .25 / .125 / .0625)
    ("swin_l_T2_B1_fusion", dict(img_size=224, num_frames=2, label_dim=29, **O.PRESETS["swin_l"]), 1, "fusion"),
]
# every case uses bf16-representable weights (make_state docstring): the same fixture then serves the fp32
# check mode (1e-3) and the bf16 production mode (2e-2) with IDENTICAL weights on both sides
STATE_KW = dict(bf16_exact=True)


def build_reference(cfg, mode):
    ref = shim.load_reference()
    model = ref.SwinTransformer2D_Adapter_New(
        label_dim=cfg.label_dim, pretrained=None, img_size=cfg.img_size, patch_size=[1, cfg.patch, cfg.patch],
        num_frames=cfg.num_frames, embed_dim=cfg.embed_dim, depths=cfg.depths, num_heads=cfg.num_heads,
        window_size=cfg.window_size, ftmode=mode, adapter_mlp_ratio=cfg.adapter_mlp_ratio, drop_path_rate=0.0)
    return model


def reference_step(model, sd, a, v, labels, mode):
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()) or set(ref_sd.keys()) == set(sd.keys()), (
        sorted(set(ref_sd) ^ set(sd)))
    for k, t in ref_sd.items():
        assert tuple(t.shape) == tuple(sd[k].shape), (k, t.shape, sd[k].shape)
        if O.is_buffer(k):  # integer / mask buffers: the reference's own values must equal ours bit-exactly
            assert torch.equal(t.to(sd[k].dtype), sd[k]), "buffer mismatch: " + k
    model.load_state_dict(sd, strict=True)
    model.train()
    model.mlp_head[1].eval()
    for name, p in model.named_parameters():
        p.requires_grad_(O.is_trainable(name))
    out = model(a, v, mode)
    loss = torch.nn.CrossEntropyLoss()(out, labels)
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters() if p.requires_grad and p.grad is not None}
    return out.detach(), loss.detach(), grads


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    if not shim.reference_available():
        print("reference not available; nothing to validate")
        return 0
    worst = 0.0
    for name, kw, batch, mode in CASES:
        cfg = O.SwinAVEConfig(**kw)
        sd = O.make_state(cfg, seed=3, dtype=torch.float64, **STATE_KW)
        a, v, labels = O.make_inputs(cfg, batch, seed=3, dtype=torch.float64)
        t0 = time.time()
        model = build_reference(cfg, mode).double()
        r_out, r_loss, r_g = reference_step(model, sd, a, v, labels, mode)
        t1 = time.time()
        o_out, o_loss, o_g = O.train_step(sd, a, v, labels, cfg, mode)
        t2 = time.time()
        e_out = rel(o_out, r_out)
        assert set(o_g.keys()) == set(r_g.keys()), sorted(set(o_g) ^ set(r_g))
        e_g = max(rel(o_g[k], r_g[k]) for k in r_g)
        n_zero = sum(1 for k in r_g if float(r_g[k].abs().max()) == 0.0)
        print(f"{name:26s} logits rel {e_out:.2e}  loss diff {abs(float(o_loss - r_loss)):.2e}  "
              f"max grad rel {e_g:.2e} over {len(r_g)} tensors ({n_zero} all-zero)  "
              f"ref {t1 - t0:.1f}s oracle {t2 - t1:.1f}s")
        worst = max(worst, e_out, e_g)
    ok = worst < 1e-9
    print("ORACLE PINNED TO REFERENCE (fp64): worst rel diff %.2e -> %s" % (worst, "OK" if ok else "MISMATCH"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
