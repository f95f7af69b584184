#!/usr/bin/env python
"""Micro-benchmark of stgcma_gemm on the step's dominant shapes (CUDA events, L2 flushed between runs by
rotating through several operand sets).  `--only i` runs a single shape once (for ncu --set full)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stg-cma_b200")]
import torch  # noqa: E402
import stgcma.lib as lib  # noqa: E402

SHAPES = [  # M, N, K, groups, epi, opts
    (1003520, 288, 96, 1, lib.EPI_LINEAR, dict(bias=True)),
    (1003520, 384, 96, 1, lib.EPI_GELU, dict(bias=True, out_pre=True)),
    (1003520, 384, 96, 1, lib.EPI_DGELU, dict(aux=True)),
    (1003520, 96, 384, 1, lib.EPI_LINEAR, dict(bias=True)),
    (1003520, 96, 96, 1, lib.EPI_LINEAR, dict(bias=True)),
    (1003520, 96, 16, 2, lib.EPI_LINEAR, dict(bias=True, res1=True, res2=True, f32=True)),
    (1003520, 16, 96, 2, lib.EPI_GELU, dict(bias=True, out_pre=True)),
    (62720, 1536, 384, 1, lib.EPI_GELU, dict(bias=True, out_pre=True)),
    (62720, 384, 1536, 1, lib.EPI_LINEAR, dict(bias=True)),
    (62720, 1152, 384, 1, lib.EPI_LINEAR, dict(bias=True)),
    (15680, 768, 3072, 1, lib.EPI_LINEAR, dict(bias=True)),
    (8192, 8192, 8192, 1, lib.EPI_LINEAR, dict()),
    (31360, 2304, 768, 1, lib.EPI_LINEAR, dict(bias=True)),
    (31360, 1152, 768, 1, lib.EPI_LINEAR, dict(bias=True)),
    (31360, 768, 768, 1, lib.EPI_LINEAR, dict(bias=True)),
    (3920, 2304, 768, 1, lib.EPI_LINEAR, dict(bias=True)),
    (31360, 2304, 384, 1, lib.EPI_LINEAR, dict(bias=True)),
    (31360, 2304, 768, 1, lib.EPI_GELU, dict(bias=True, out_pre=True)),
]


def run(i, reps, sets):
    M, N, K, g, epi, o = SHAPES[i]
    dev = "cuda"
    bf = torch.bfloat16
    odt = torch.float32 if o.get("f32") else bf
    ops = []
    for _ in range(sets):
        ops.append(dict(A=torch.randn(M, K, device=dev, dtype=bf), W=torch.randn(g, N, K, device=dev, dtype=bf) * K ** -0.5,
                        out=torch.empty(M, N, device=dev, dtype=odt),
                        bias=torch.randn(g, N, device=dev) if o.get("bias") else None,
                        res1=torch.randn(M, N, device=dev, dtype=odt) if o.get("res1") else None,
                        res2=torch.randn(M, N, device=dev, dtype=bf) if o.get("res2") else None,
                        aux=torch.randn(M, N, device=dev, dtype=bf) if o.get("aux") else None,
                        out_pre=torch.empty(M, N, device=dev, dtype=bf) if o.get("out_pre") else None))

    def call(d):
        lib.gemm(d["A"], d["W"], d["out"], M, N, K, bias=d["bias"], res1=d["res1"], res2=d["res2"], aux=d["aux"],
                 out_pre=d["out_pre"], groups=g, epi=epi)

    for d in ops:
        call(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        call(ops[r % sets])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    es = 2
    nb = es * (M * K + g * N * K) + M * N * ((4 if o.get("f32") else 2) * (1 + bool(o.get("res1"))) + es * (
        bool(o.get("res2")) + bool(o.get("aux")) + bool(o.get("out_pre"))))
    print(f"{i:2d} M={M} N={N} K={K} g={g} epi={epi} {sorted(o)}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s  "
          f"{nb / ms / 1e6:6.0f} GB/s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--from", dest="first", type=int, default=0)
    args = ap.parse_args()
    if args.only >= 0:
        run(args.only, 2, 1)
        return
    for i in range(args.first, len(SHAPES)):
        M, N, K = SHAPES[i][:3]
        sets = max(1, min(6, int(3e8 // (M * (N + K) * 2)) + 1))   # rotate operand sets when one set fits in L2
        run(i, args.reps, sets)


if __name__ == "__main__":
    main()
