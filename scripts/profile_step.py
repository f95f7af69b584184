#!/usr/bin/env python
"""One adapter train step bracketed by cudaProfilerStart/Stop, for ncu (never a bench number):

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python scripts/profile_step.py --batch 4
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc -c 3 \
        -o gpurun_out/prof_gemm python scripts/profile_step.py --batch 4
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stg-cma_b200")]
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="swin_t")
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--warm", type=int, default=2)
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

    m.engine().side_stream_dw = False   # one chain on one stream: the launch list is the whole step, serialised
    m.engine().graph_mode = "0"         # eager launches (ncu lists them the same either way; keeps the launch order explicit)
    m.engine().lanes = 1
    for _ in range(args.warm):
        step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
