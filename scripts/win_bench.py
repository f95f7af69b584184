#!/usr/bin/env python
"""Times stgcma_window_attn_fwd / _bwd on the Swin-T shapes at 64 clips per GPU (1280 frames = video + audio x 10) against the
HBM floor of the call (fwd: qkv read + out written = 4 units; bwd: qkv + dO read, dqkv written = 7 units of 2 * rows * C bytes).

    python scripts/win_bench.py [--frames 1280]
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
    ap.add_argument("--frames", type=int, default=1280)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--hbm", type=float, default=6459.9)
    args = ap.parse_args()
    for R, heads in ((56, 3), (28, 6), (14, 12), (7, 24)):
        for shift in (0, 3):
            if R == 7 and shift:
                continue
            C = heads * 32
            rows = args.frames * R * R
            qkv = torch.randn(rows, 3 * C, device="cuda").to(torch.bfloat16)
            rpb = torch.randn(169, heads, device="cuda") * 0.5
            out, dout, dqkv = torch.empty(rows, C, device="cuda", dtype=torch.bfloat16), torch.randn(rows, C, device="cuda").to(torch.bfloat16), torch.empty_like(qkv)
            res = {}
            for what, fn in (("fwd", lambda: L.window_attn_fwd(qkv, rpb, out, args.frames, R, 7, shift, heads)),
                             ("bwd", lambda: L.window_attn_bwd(qkv, rpb, dout, dqkv, args.frames, R, 7, shift, heads))):
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
            unit = 2.0 * rows * C
            print(f"R {R:2d} heads {heads:2d} shift {shift}: fwd {res['fwd']:.3f} ms = {4 * unit / res['fwd'] / 1e6:.0f} GB/s ({4 * unit / res['fwd'] / 1e6 / args.hbm:.2f} of HBM)   "
                  f"bwd {res['bwd']:.3f} ms = {7 * unit / res['bwd'] / 1e6:.0f} GB/s ({7 * unit / res['bwd'] / 1e6 / args.hbm:.2f})")


if __name__ == "__main__":
    main()
