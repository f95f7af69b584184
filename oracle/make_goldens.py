"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.pt from the UNMODIFIED reference.

The reference ships no golden vectors, so the fixtures are outputs of the
reference module itself (imported via oracle/shim.py), produced in the build container:

    python -m oracle.make_goldens

Each fixture stores the case definition (config kwargs, seeds, mode) -- the weights and
inputs are re-created from the seeds by oracle.swin_ave_oracle.make_state / make_inputs --
plus the reference's fp64 logits, loss, and gradients (all of them for the tiny case; for
the full-width Swin-T case the small tensors in full and an L2 norm + 64-sample probe of
every tensor, to keep the fixture small).
"""
import os
import sys

import torch

from . import shim
from . import swin_ave_oracle as O
from .validate_oracle import CASES, STATE_KW, build_reference, reference_step

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 3


def probe_index(numel, n=64):
    g = torch.Generator().manual_seed(numel)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def main():
    assert shim.reference_available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    only = set(sys.argv[1:])   # optional: regenerate only the named cases
    for name, kw, batch, mode in CASES:
        if only and name not in only:
            continue
        cfg = O.SwinAVEConfig(**kw)
        sd = O.make_state(cfg, seed=SEED, dtype=torch.float64, **STATE_KW)
        a, v, labels = O.make_inputs(cfg, batch, seed=SEED, dtype=torch.float64)
        model = build_reference(cfg, mode).double()
        out, loss, grads = reference_step(model, sd, a, v, labels, mode)
        full = name.startswith("tiny")
        fx = dict(name=name, cfg=kw, batch=batch, mode=mode, seed=SEED, state_kw=dict(STATE_KW),
                  logits=out, loss=loss,
                  state_keys=list(sd.keys()),
                  state_shapes={k: tuple(t.shape) for k, t in sd.items()},
                  grad_norm={k: float(g.norm()) for k, g in grads.items()},
                  grad_probe={k: g.reshape(-1)[probe_index(g.numel())].clone() for k, g in grads.items()},
                  grads={k: g.clone() for k, g in grads.items() if full or g.numel() <= 4096})
        path = os.path.join(OUT, name + ".pt")
        torch.save(fx, path)
        print(f"wrote {path}: {os.path.getsize(path) / 1024:.0f} KiB, {len(grads)} grads, "
              f"{len(fx['grads'])} stored in full")


if __name__ == "__main__":
    main()
