"""TEST INFRASTRUCTURE ONLY.  Pins oracle/clip_ave_oracle.py to the UNMODIFIED reference `MM_CLIP_AVE`
(AVE/model/CLIP_AVE.py:716) in fp64, exactly like oracle/validate_oracle.py does for the Swin model: strict
state_dict load of the oracle's seeded state (names + shapes), the caller's freeze rule, forward + soft-target
CrossEntropy + backward on both sides, max |diff| / max |ref| over logits and every trainable gradient.

    python -m oracle.validate_clip_oracle       # build container only (/root/reference)
"""
import sys
import time

import torch

from . import shim
from . import clip_ave_oracle as C

CASES = [
    # name, cfg kwargs, batch, mode
    ("clip_tiny_fusion", dict(label_dim=7, input_resolution=64, audio_length=480, num_frames=3, patch_size=16, embed_dim=64,
                              layers=2, heads=4), 2, "fusion"),
    ("clip_tiny_multimodal", dict(label_dim=7, input_resolution=64, audio_length=480, num_frames=3, patch_size=16, embed_dim=64,
                                  layers=2, heads=4), 2, "multimodal"),
    # full-width ViT-B/16 geometry (197 video / 49 audio tokens, D = 768, 8 heads, r = 48), 2 layers, 2 frames, 1 clip
    ("clip_b16_L2_T2_fusion", dict(label_dim=29, input_resolution=224, audio_length=1024, num_frames=2, patch_size=16,
                                   embed_dim=768, layers=2, heads=8), 1, "fusion"),
]
STATE_KW = dict(bf16_exact=True)


def build_reference(cfg, mode):
    ref = shim.load_reference("AVE/model/CLIP_AVE.py", "_ref_CLIP_AVE")
    return ref.MM_CLIP_AVE(label_dim=cfg.label_dim, input_resolution=cfg.input_resolution, audio_length=cfg.audio_length,
                           num_video_frames=cfg.num_frames, patch_size=cfg.patch_size, embed_dim=cfg.embed_dim,
                           layers=cfg.layers, heads=cfg.heads, drop_path_rate=0.0, pretrained=None, ftmode=mode)


def reference_step(model, sd, a, v, labels, mode):
    ref_sd = model.state_dict()
    assert set(ref_sd.keys()) == set(sd.keys()), sorted(set(ref_sd) ^ set(sd))
    for k, t in ref_sd.items():
        assert tuple(t.shape) == tuple(sd[k].shape), (k, t.shape, sd[k].shape)
    model.load_state_dict(sd, strict=True)     # (copies into the fp32 LayerNorm parameters with a cast)
    model.train()
    model.mlp_head[1].eval()
    for name, p in model.named_parameters():
        p.requires_grad_(C.is_trainable(name))
    out = model(a, v, mode)
    loss = torch.nn.CrossEntropyLoss()(out, labels)
    loss.backward()
    return out.detach(), loss.detach(), {n: p.grad for n, p in model.named_parameters() if p.requires_grad and p.grad is not None}


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def main():
    if not shim.reference_available():
        print("reference not available; nothing to validate")
        return 0
    worst = 0.0
    for name, kw, batch, mode in CASES:
        cfg = C.ClipAVEConfig(**kw)
        sd = C.make_state(cfg, seed=3, dtype=torch.float64, **STATE_KW)
        a, v, labels = C.make_inputs(cfg, batch, seed=3, dtype=torch.float64)
        t0 = time.time()
        model = build_reference(cfg, mode).double()
        ref = shim.load_reference("AVE/model/CLIP_AVE.py", "_ref_CLIP_AVE")
        for mod in model.modules():          # the reference LayerNorm casts its INPUT to fp32 (:38): its parameters must be fp32
            if isinstance(mod, ref.LayerNorm):
                mod.float()
        C.LN_FP32 = True
        r_out, r_loss, r_g = reference_step(model, sd, a, v, labels, mode)
        t1 = time.time()
        o_out, o_loss, o_g = C.train_step(sd, a, v, labels, cfg, mode)
        t2 = time.time()
        e_out = rel(o_out, r_out)
        assert set(o_g.keys()) == set(r_g.keys()), sorted(set(o_g) ^ set(r_g))
        e_g = max(rel(o_g[k], r_g[k]) for k in r_g)
        n_zero = sum(1 for k in r_g if float(r_g[k].abs().max()) == 0.0)
        print(f"{name:26s} logits rel {e_out:.2e}  loss diff {abs(float(o_loss - r_loss)):.2e}  "
              f"max grad rel {e_g:.2e} over {len(r_g)} tensors ({n_zero} all-zero)  ref {t1 - t0:.1f}s oracle {t2 - t1:.1f}s")
        worst = max(worst, e_out, e_g)
    # logits agree exactly; the gradients carry the fp32 round-off of the reference's forced-fp32 LayerNorm (autograd orders the
    # casts differently on the two sides): 5e-8, against 1e-14 for the Swin model whose LayerNorms run in fp64
    ok = worst < 1e-6
    print("CLIP ORACLE PINNED TO REFERENCE (fp64): worst rel diff %.2e -> %s" % (worst, "OK" if ok else "MISMATCH"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
