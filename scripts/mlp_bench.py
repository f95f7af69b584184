#!/usr/bin/env python
"""Micro-benchmark of the fused MLP kernels (stgcma_mlp_fwd / _bwd) on the Swin-T stage shapes at B clips per GPU
(CUDA events; operand sets rotated so that inputs do not stay in L2).  --trace dumps CTA 0's phase timeline of the
forward kernel (debug hook stgcma_debug_trace)."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stg-cma_b200")]
import torch  # noqa: E402
import stgcma.lib as lib  # noqa: E402


def run(M, C, rp, reps, sets, which, trace):
    H = 4 * C
    dev, bf = "cuda", torch.bfloat16
    ops = []
    for _ in range(sets):
        d = dict(x=torch.randn(M, C, device=dev), y=torch.empty(M, C, device=dev, dtype=bf),
                 hpre=torch.empty(M, rp, device=dev, dtype=bf), h=torch.empty(M, rp, device=dev, dtype=bf),
                 mean=torch.empty(M, device=dev), rstd=torch.empty(M, device=dev),
                 dhpre=torch.randn(M, rp, device=dev, dtype=bf), dxb=torch.randn(M, C, device=dev, dtype=bf),
                 dxn=torch.empty(M, C, device=dev, dtype=bf))
        ops.append(d)
    g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    W1 = torch.randn(H, C, device=dev, dtype=bf) * C ** -0.5
    W2 = torch.randn(C, H, device=dev, dtype=bf) * H ** -0.5
    b1, b2 = torch.zeros(H, device=dev), torch.zeros(C, device=dev)
    Wa = torch.randn(2, rp, C, device=dev, dtype=bf) * C ** -0.5
    ba = torch.zeros(2, rp, device=dev)
    Wat, W2t, W1t = Wa.transpose(1, 2).contiguous(), W2.t().contiguous(), W1.t().contiguous()

    def fwd(d):
        lib.mlp_fwd(d["x"], g, b, W1, b1, W2, b2, Wa, ba, d["y"], d["hpre"], d["h"], d["mean"], d["rstd"], M, C, H, rp)

    def bwd(d):
        lib.mlp_bwd(d["dhpre"], d["dxb"], d["x"], g, b, d["mean"], d["rstd"], Wat, W1, b1, W2t, W1t, d["dxn"], M, C, H, rp)

    fn = fwd if which == "fwd" else bwd
    for d in ops:
        fwd(d)
        fn(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        fn(ops[r % sets])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * M * (2 * C * H + C * rp) * (1.5 if which == "bwd" else 1.0)   # bwd executes the fc1 recompute too
    nbytes = M * (4 * C + 2 * C + 4 * rp + 8) if which == "fwd" else M * (4 * C + 2 * C + 2 * C + 2 * rp + 8)
    tiles = 2 * ((M // 2 + 127) // 128)
    print(f"mlp_{which} M={M} C={C} rp={rp}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TFLOP/s (executed)  "
          f"{nbytes / ms / 1e6:7.1f} GB/s (algorithmic)  {ms * 1e3 / (tiles / 148):6.2f} us per tile per SM", flush=True)
    if trace and which == "fwd":
        buf = torch.zeros(3 * 512, dtype=torch.int64, device=dev)
        h = lib.load()
        h.stgcma_debug_trace.argtypes = [ctypes.c_void_p]
        h.stgcma_debug_trace(buf.data_ptr())
        fwd(ops[0])
        torch.cuda.synchronize()
        h.stgcma_debug_trace(None)
        t = buf.cpu().view(3, 512)
        for role, name in ((0, "worker0"), (1, "mma"), (2, "tma")):
            ev = [(int(v) >> 48, int(v) & 0xffffffffffff) for v in t[role].tolist() if v != 0]
            if not ev:
                continue
            t0 = ev[0][1]
            print(f"  trace {name}: " + " ".join(f"{i}@{(c - t0)}" for i, c in ev[:160]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--sets", type=int, default=3)
    ap.add_argument("--trace", action="store_true")
    args = ap.parse_args()
    rows0 = 2 * args.batch * 10 * 3136
    for which in ("fwd", "bwd"):
        run(rows0, 96, 16, args.reps, args.sets, which, args.trace)
        run(rows0 // 4, 192, 32, args.reps, args.sets, which, args.trace)


if __name__ == "__main__":
    main()
