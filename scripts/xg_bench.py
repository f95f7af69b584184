#!/usr/bin/env python
"""Times stgcma_xmodal_global_fwd / _bwd on the Swin-T shapes of the global cross-modal attention at 64 clips per GPU
(640 frame pairs): stage 0 (L = 3136, r = 12 -> rp 16), stage 1 (L = 784, rp 32), stage 2 (L = 196, rp 32).

    python scripts/xg_bench.py [--frames 640]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stg-cma_b200"))
import torch  # noqa: E402
import stgcma.lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=640)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    for Ln, rp, r in ((3136, 16, 12), (784, 32, 24), (196, 32, 24)):
        rows = 2 * args.frames * Ln
        h = (torch.randn(rows, rp, device="cuda") * 0.5).to(torch.bfloat16)
        h[:, r:] = 0
        hpre = torch.randn(rows, rp, device="cuda").to(torch.bfloat16)
        gate = torch.tensor([0.7, 0.3], device="cuda")
        out, att, dh = torch.empty_like(h), torch.empty_like(h), torch.empty_like(h)
        lse, work = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
        dout = torch.randn(rows, rp, device="cuda").to(torch.bfloat16)
        dgate = torch.zeros(2, device="cuda")
        res = {}
        for what, fn in (("fwd", lambda: L.xmodal_global_fwd(h, gate, out, att, lse, args.frames, Ln, rp, r=r)),
                         ("bwd", lambda: L.xmodal_global_bwd(h, hpre, gate, att, lse, dout, dh, dgate, work, args.frames, Ln, rp, r=r))):
            for _ in range(2):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[what] = e0.elapsed_time(e1) / args.reps
        ex = 2.0 * args.frames * Ln * Ln          # exponentials of one forward (both directions)
        floor = ex / (148 * 16 * 1.9e9) * 1e3
        print(f"L {Ln:5d} rp {rp}: fwd {res['fwd']:.3f} ms (MUFU-only floor {floor:.3f} ms, {floor / res['fwd']:.2f})   "
              f"bwd {res['bwd']:.3f} ms ({2 * floor / res['bwd']:.2f} of its 2-exp floor)")


if __name__ == "__main__":
    main()
