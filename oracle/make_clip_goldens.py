"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/clip_*.pt from the UNMODIFIED reference `MM_CLIP_AVE`
(build container only):

    python -m oracle.make_clip_goldens

Same layout as oracle/make_goldens.py: the case definition (config kwargs, seeds, mode; weights and inputs are re-created
from the seeds by oracle.clip_ave_oracle.make_state / make_inputs) plus the reference's logits, loss and gradients -- all
of them for the tiny cases; for the full-width ViT-B/16 case the small tensors in full and an L2 norm + 64-sample probe of
every tensor.
"""
import os

import torch

from . import shim
from . import clip_ave_oracle as C
from .make_goldens import probe_index
from .validate_clip_oracle import CASES, STATE_KW, build_reference, reference_step

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 3


def main():
    assert shim.reference_available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    ref = shim.load_reference("AVE/model/CLIP_AVE.py", "_ref_CLIP_AVE")
    for name, kw, batch, mode in CASES:
        cfg = C.ClipAVEConfig(**kw)
        sd = C.make_state(cfg, seed=SEED, dtype=torch.float64, **STATE_KW)
        a, v, labels = C.make_inputs(cfg, batch, seed=SEED, dtype=torch.float64)
        model = build_reference(cfg, mode).double()
        for mod in model.modules():
            if isinstance(mod, ref.LayerNorm):
                mod.float()
        out, loss, grads = reference_step(model, sd, a, v, labels, mode)
        full = "tiny" in name
        fx = dict(name=name, cfg=kw, batch=batch, mode=mode, seed=SEED, state_kw=dict(STATE_KW), logits=out, loss=loss,
                  state_keys=list(sd.keys()), state_shapes={k: tuple(t.shape) for k, t in sd.items()},
                  grad_norm={k: float(g.norm()) for k, g in grads.items()},
                  grad_probe={k: g.reshape(-1)[probe_index(g.numel())].clone() for k, g in grads.items()},
                  grads={k: g.clone() for k, g in grads.items() if full or g.numel() <= 4096})
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(grads)} grads, {len(fx['grads'])} stored in full")


if __name__ == "__main__":
    main()
