#!/usr/bin/env python
"""Per-shape timing of every stgcma_gemm / gemm_tn call in one train step (CUDA events around each call).
    python scripts/gemm_shapes.py --batch 16
Prints calls, total ms, TFLOP/s and algorithmic GB/s per (M, N, K, groups, epilogue) so the GEMM work can be
judged shape by shape against both rooflines (small-K shapes are HBM-bound, not tensor-bound)."""
import argparse
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stg-cma_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
import stgcma.lib as lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="swin_t")
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args()
    from stgcma import synth
    dev = torch.device("cuda", 0)
    m, opt = bench.build_model(args.model, torch.bfloat16, dev)
    a, v, y = (t.to(dev) for t in synth.make_clips(m.cfg, args.batch, seed=0, pin=False))
    loss_fn = torch.nn.CrossEntropyLoss()

    def step():
        out = m(a, v, "fusion")
        loss = loss_fn(out, y)
        opt.zero_grad()
        loss.backward()
        opt.step()

    for _ in range(2):
        step()
    rec = []
    orig_gemm, orig_tn = lib.gemm, lib.gemm_tn

    def gemm(A, W, out, M, N, K, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_gemm(A, W, out, M, N, K, **kw)
        e1.record()
        es = A.element_size()
        nb = es * (M * K + kw.get("groups", 1) * N * K) + M * N * (out.element_size() * (1 + (kw.get("res1") is not None)) + es * (
            (kw.get("res2") is not None) + (kw.get("aux") is not None) + (kw.get("out_pre") is not None)))
        key = ("gemm", str(A.dtype)[6:], M, N, K, kw.get("groups", 1), kw.get("epi", 0),
               "r1" * (kw.get("res1") is not None) + "r2" * (kw.get("res2") is not None) + "pre" * (kw.get("out_pre") is not None)
               + "aux" * (kw.get("aux") is not None) + ("f32o" if out.dtype == torch.float32 and A.dtype != torch.float32 else ""))
        rec.append((key, e0, e1, 2.0 * M * N * K, nb))

    def gemm_tn(A, B, dW, M, Ka, Nb, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_tn(A, B, dW, M, Ka, Nb, **kw)
        e1.record()
        rec.append((("gemm_tn", str(A.dtype)[6:], M, Ka, Nb, 1, 0, ""), e0, e1, 2.0 * M * Ka * Nb, A.element_size() * M * (Ka + Nb)))

    import stgcma.engine as E
    E.L_.gemm, E.L_.gemm_tn = gemm, gemm_tn
    step()
    torch.cuda.synchronize()
    agg = OrderedDict()
    for key, e0, e1, fl, nb in rec:
        d = agg.setdefault(key, [0, 0.0, 0.0, 0.0, 1e9, 0.0])
        ms = e0.elapsed_time(e1)
        d[0] += 1
        d[1] += ms
        d[4], d[5] = min(d[4], ms), max(d[5], ms)
        d[2] += fl
        d[3] += nb
    tot = sum(d[1] for d in agg.values())
    print(f"batch {args.batch}: {len(rec)} calls, {tot:.2f} ms total")
    print("| kind | dtype | M | N(Ka) | K(Nb) | g | epi | flags | calls | ms | TFLOP/s | GB/s | min us | max us |")
    print("|---|---|---:|---:|---:|---:|---:|---|---:|---:|---:|---:|---:|---:|")
    for key, d in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| " + " | ".join(str(k) for k in key) + f" | {d[0]} | {d[1]:.3f} | {d[2] / d[1] / 1e9:.1f} | {d[3] / d[1] / 1e6:.0f} | {d[4] * 1e3:.0f} | {d[5] * 1e3:.0f} |")


if __name__ == "__main__":
    main()
