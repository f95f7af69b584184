#!/usr/bin/env python
"""Times stgcma_mha_fwd / stgcma_mha_bwd on the four attention shapes of one CLIP ViT-B/16 block at 64 clips per GPU
(spatial 197 / 49 tokens, temporal T = 10 over 197 / 49 token positions; 8 heads of 96):

    python scripts/mha_bench.py [--clips 64] [--hd 96] [--heads 8]
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
    ap.add_argument("--clips", type=int, default=64)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--hd", type=int, default=96)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    B, T, heads, hd = args.clips, 10, args.heads, args.hd
    D = heads * hd
    shapes = [("spatial video", B * T, 197, 1), ("spatial audio", B * T, 49, 1), ("temporal video", B * 197, T, 197),
              ("temporal audio", B * 49, T, 49)]
    for name, n_seq, seq_len, inner in shapes:
        rows = n_seq * seq_len
        qkv = torch.randn(rows, 3 * D, device="cuda", dtype=torch.bfloat16)
        out = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(rows, heads, device="cuda")
        dout = torch.randn(rows, D, device="cuda", dtype=torch.bfloat16)
        dqkv = torch.empty_like(qkv)
        work = torch.empty(rows, heads, device="cuda")
        res = {}
        for what, fn in (("fwd", lambda: L.mha_fwd(qkv, out, lse, n_seq, seq_len, inner, heads, hd)),
                         ("bwd", lambda: L.mha_bwd(qkv, out, lse, dout, dqkv, work, n_seq, seq_len, inner, heads, hd))):
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
        fl = 4.0 * rows * seq_len * D
        by = rows * D * 2
        print(f"{name:15s} n_seq {n_seq:6d} x {seq_len:3d}: fwd {res['fwd']:.3f} ms ({fl / res['fwd'] / 1e9:.0f} TF/s, {4 * by / res['fwd'] / 1e6:.0f} GB/s)"
              f"   bwd {res['bwd']:.3f} ms ({2.5 * fl / res['bwd'] / 1e9:.0f} TF/s algorithmic, {8 * by / res['bwd'] / 1e6:.0f} GB/s)")


if __name__ == "__main__":
    main()
