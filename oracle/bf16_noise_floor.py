"""TEST INFRASTRUCTURE ONLY (analysis).  What does bf16 cost the REFERENCE itself?

Runs the UNMODIFIED reference module (through oracle/shim.py) on the tiny parity configuration twice -- in fp64, and in fp32
parameters under torch.autocast(bfloat16) (the mode the caller trains in, traintest_adapt_ave29.py:158-162) -- and reports the
per-tensor gradient error of the autocast run against the fp64 run, for several seeds.  This is the noise floor that the bf16
production mode of this repository is judged against: if the reference's own bf16 run differs from its fp64 run by e on a tensor,
no bf16 implementation can be expected to sit much below e on it.

    python -m oracle.bf16_noise_floor        (needs /root/reference; CPU, ~1 min)
"""
import sys

import torch

from . import shim
from . import swin_ave_oracle as O
from .validate_oracle import build_reference

KW = dict(img_size=112, num_frames=3, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
          adapter_mlp_ratio=[0.25, 0.25, 0.125], label_dim=7)


def step(model, sd, a, v, labels, autocast):
    model.load_state_dict(sd, strict=True)
    model.train()
    model.mlp_head[1].eval()
    for n, p in model.named_parameters():
        p.requires_grad_(O.is_trainable(n))
        p.grad = None
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        out = model(a, v, "fusion")
        loss = torch.nn.CrossEntropyLoss()(out.float() if autocast else out, labels)
    loss.backward()
    return out.detach().double(), {n: p.grad.double() for n, p in model.named_parameters() if p.requires_grad and p.grad is not None}


def main():
    if not shim.reference_available():
        print("reference not available")
        return 0
    cfg = O.SwinAVEConfig(**KW)
    for seed in (3, 5, 11):
        sd = O.make_state(cfg, seed=seed, bf16_exact=True)
        a, v, labels = O.make_inputs(cfg, 2, seed=seed)
        m64 = build_reference(cfg, "fusion").double()
        o64, g64 = step(m64, {k: (t.double() if t.is_floating_point() else t) for k, t in sd.items()}, a.double(), v.double(), labels.double(), False)
        m32 = build_reference(cfg, "fusion")
        o16, g16 = step(m32, sd, a, v, labels, True)
        rows = []
        for n, r in g64.items():
            if "gate_" in n:
                continue
            g = g16[n]
            rows.append((float((g - r).norm() / r.norm().clamp_min(1e-30)), float((g - r).abs().max() / r.abs().max().clamp_min(1e-30)), n))
        rows.sort(reverse=True)
        tt = [r for r in rows if "temporal_position_bias_table" in r[2]]
        print(f"seed {seed}: logits {float((o16 - o64).abs().max() / o64.abs().max()):.2e}; temporal bias tables: worst rel-L2 {max(r[0] for r in tt):.2e}, "
              f"worst max-abs {max(r[1] for r in tt):.2e}; all tensors: worst rel-L2 {rows[0][0]:.2e} ({rows[0][2]}), median {rows[len(rows) // 2][0]:.2e}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
