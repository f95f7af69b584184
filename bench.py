#!/usr/bin/env python
"""bench.py -- adapter-train clips/sec of the AVE Swin cross-modal adapter step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--model swin_t] [--batch B]      # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W                  # the reference's CPU path

One "step" = the caller's loop body (AVE/traintest_adapt_ave29.py:158-168): forward of
`SwinTransformer2D_Adapter_New(a, v, 'fusion')`, soft-target CrossEntropy, backward (frozen backbone: dX
everywhere, dW for adapters / gates / temporal tables / head), the gradient-bucket all-reduce when N > 1,
and the caller's Adam step.  Workload: BASELINE.json configs[0]/[2] -- AVE 29-class Swin-T
(embed 96, depths [2,2,6,2], heads [3,6,12,24], adapter ratios [.125,.125,.0625,.0625]), 10 frames 3x224^2
+ 10 audio 224^2 spectrogram images per clip, bf16 storage / fp32 accumulate, synthetic data, seeded
random weights.  N > 1: one process per GPU (torchrun), clips sharded across ranks (weak scaling:
`--batch` clips per GPU), no data-path collective except the adapter-gradient bucket.

`value`  : clips/s with the step's inputs already resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric through the public module call with HOST (pinned) inputs: H2D copies of the clips
           and labels and the D2H read of the loss are inside the timed region.
`roofline`: the kernel class with the largest share of step time, measured live by bracketing every
           C-ABI launch with CUDA events in one extra profiled step; achieved = algorithmic FLOPs (or
           bytes) of those launches / their summed duration; peak from MEASURED_PEAKS.json.
`cpu_baseline`: the oracle (a torch-CPU port of the reference's algorithm, pinned to the reference in
           fp64) timed on this box's host cores on a bounded sample (kind "port": the Python reference
           itself cannot travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "stg-cma_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "adapter_train_clips_per_sec"
UNIT = "clips/s"
# algorithmic FLOPs per clip, forward + frozen-aware backward (BASELINE.md section 2 / SURVEY.md 8d)
FLOPS_PER_CLIP = {"swin_t": 501.6e9, "swin_b": 1587.5e9, "swin_l": 3989.0e9}
WORKLOAD = {"swin_t": "AVE-29 Swin-T fusion adapter train step (BASELINE.json configs[0]/[2])",
            "swin_b": "AVE-29 Swin-B fusion adapter train step (run_adapt_ave29.py MM-Swin-AVE-Base)",
            "swin_l": "AVE-29 Swin-L fusion adapter train step (run_adapt_ave29.py MM-Swin-AVE-Large)"}


def measured_traffic(kernel, batch, model):
    """DRAM bytes per launch of a kernel class from the newest committed ncu traffic summary for this workload
    (scripts/traffic_summary.py); None when there is none for this model / batch."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"*_traffic_{model}_b{batch}.json")), reverse=True):
        try:
            d = json.load(open(path))["classes"].get(kernel)
            if d:
                return d["dram_bytes_per_launch"], os.path.relpath(path, ROOT)
        except (OSError, ValueError, KeyError):
            continue
    return None, None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_model(name, compute_dtype, device):
    import model
    from stgcma.config import PRESETS
    from stgcma import synth
    torch.manual_seed(0)
    m = model.Swin_AVE.SwinTransformer2D_Adapter_New(label_dim=29, patch_size=[1, 4, 4], num_frames=10, window_size=7,
                                                     pretrained=None, ftmode="fusion", compute_dtype=compute_dtype,
                                                     **PRESETS[name])
    synth.randomize_(m, seed=0)
    m = m.to(device)
    head, adapt, base = synth.split_parameters(m)
    for p in base:                       # freeze_base=True (traintest_adapt_ave29.py:58-61)
        p.requires_grad_(False)
    # traintest_adapt_ave29.py:68 (head_lr = 0.1 per the Swin launch script); FusedAdam = torch.optim.Adam's update rule
    # and state layout as one kernel launch over all ~455 trainable tensors (stgcma_adam_step)
    from stgcma.optim import FusedAdam
    opt = FusedAdam([{"params": adapt, "lr": 5e-5}, {"params": head, "lr": 5e-6}], weight_decay=5e-7, betas=(0.95, 0.999))
    m.train()
    return m, opt


def ours(args):
    import stgcma.lib as lib
    from stgcma import synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    m, opt = build_model(args.model, torch.bfloat16, dev)
    cfg = m.cfg
    loss_fn = torch.nn.CrossEntropyLoss()
    B = args.batch
    a_h, v_h, y_h = synth.make_clips(cfg, B, seed=rank, pin=True)
    a_d, v_d, y_d = a_h.to(dev), v_h.to(dev), y_h.to(dev)

    def step_resident():
        out = m(a_d, v_d, "fusion")
        loss = loss_fn(out, y_d)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss

    # e2e: the step's inputs come from pinned HOST memory every step.  The copies run on a side stream into one of two
    # device buffers (what DataLoader(pin_memory=True) + non_blocking .to() gives a training loop), so the H2D of step
    # i + 1 overlaps the compute of step i; every copy and the D2H read of the loss are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [tuple(torch.empty_like(t, device=dev) for t in (a_h, v_h, y_h)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "prefetched": False}

    def h2d(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])       # the step that last read this buffer has finished
            for dst, src in zip(bufs[slot], (a_h, v_h, y_h)):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_begin():
        for ev in consumed:
            ev.record()
        e2e_state["i"], e2e_state["prefetched"] = 0, False

    def step_e2e(last=False):
        slot = e2e_state["i"] & 1
        if not e2e_state["prefetched"]:
            h2d(slot)                                    # first step of a timed run: its own copy, not hidden
        if not last:
            h2d(slot ^ 1)                                # next step's clips, overlapping this step's compute
        e2e_state["prefetched"] = not last
        torch.cuda.current_stream().wait_event(ready[slot])
        a, v, y = bufs[slot]
        out = m(a, v, "fusion")
        loss = loss_fn(out, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        consumed[slot].record()
        e2e_state["i"] += 1
        return loss.item()           # device -> host read of the step's result (traintest_adapt_ave29.py:170)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, e2e=False):
        barrier()
        if e2e:
            e2e_begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            if e2e:
                fn(last=(i == steps - 1))    # exactly `steps` H2D copies inside the timed region
            else:
                fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.launches()
    ms = timed(step_resident, args.steps)
    launches = lib.launches() - n0
    clocks = sampler.stop() if rank == 0 else None
    e2e_begin()
    for i in range(2):
        step_e2e(last=(i == 1))
    ms_e2e = timed(step_e2e, args.steps, e2e=True)

    # one profiled step: per-kernel CUDA-event durations (not part of any reported throughput)
    # every rank runs the step (its backward holds the gradient all-reduce); only rank 0 brackets the launches
    prof = None
    torch.cuda.synchronize()
    eng = m.engine()
    # one chain on one stream for this step (no micro-batch lanes, no dW side stream): clean per-kernel durations
    side_dw, eng.side_stream_dw = eng.side_stream_dw, False
    lanes, eng.lanes = eng.lanes, 1
    if rank == 0:
        lib.profile_start()
    step_resident()
    if rank == 0:
        prof = lib.profile_stop()
    eng.side_stream_dw, eng.lanes = side_dw, lanes
    barrier()

    total_clips = B * world * args.steps
    value = total_clips / (ms / 1e3)
    e2e_value = total_clips / (ms_e2e / 1e3)
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    pk = peaks()
    tot_ms = sum(d["ms"] for d in prof.values()) or 1.0
    shares = {k.replace("stgcma_", ""): round(d["ms"] / tot_ms, 4) for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    top_name, top = max(prof.items(), key=lambda kv: kv[1]["ms"])
    # The dominant kernel class is judged against the roof it is closest to: a GEMM with K = 96 (Swin stage 0) has
    # ~70 FLOP per byte, far below the B200 ridge (~250), so its roof is HBM, not the tensor pipe.  Both fractions
    # are reported; `bound` names the larger one.
    tf = top["flops"] / (top["ms"] / 1e3) / 1e12
    gbs = top["bytes"] / (top["ms"] / 1e3) / 1e9
    f_tensor, f_hbm = tf / pk["tf_sustained"], gbs / pk["hbm"]
    if top["flops"] > 0 and f_tensor >= f_hbm:
        roof = {"kernel": top_name, "bound": "tensor", "achieved": round(tf, 2), "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": round(f_tensor, 4), "traffic": None, "hbm_frac": round(f_hbm, 4)}
    else:
        roof = {"kernel": top_name, "bound": "hbm", "achieved": round(gbs, 1), "peak": pk["hbm"], "unit": "GB/s",
                "frac": round(f_hbm, 4), "traffic": None, "tensor_frac": round(f_tensor, 4)}
    roof["peak_source"] = pk["source"] + ("; sustained figure: kernel timed inside a long step" if roof["bound"] == "tensor" else "")
    roof["traffic"], tsrc = measured_traffic(top_name, B, args.model)
    roof["algorithmic_bytes_per_launch"] = round(top["bytes"] / top["calls"])
    if tsrc:
        roof["traffic_source"] = tsrc
    roof["launches"] = top["calls"]
    roof["avg_launch_ms"] = round(top["ms"] / top["calls"], 4)
    roof["time_shares"] = shares
    roof["abi_kernel_ms_per_step"] = round(tot_ms, 3)   # sum over C-ABI launches of one profiled step (serialised by the events)
    gemm = prof.get("stgcma_gemm")
    if gemm:
        roof["gemm_tflops"] = round(gemm["flops"] / (gemm["ms"] / 1e3) / 1e12, 2)
    step_frac = value / world * FLOPS_PER_CLIP[args.model] / (pk["tf_sustained"] * 1e12)
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD[args.model], "clips_per_gpu": B, "global_batch": B * world, "frames": 10,
                   "parallelism": f"dp{world} (clips sharded, adapter-grad bucket all-reduce)" if world > 1 else "single GPU",
                   "streams": f"{m.engine().lanes} micro-batch lane(s) per GPU on separate CUDA streams + dW side stream",
                   "optimizer": "Adam (the caller's groups / betas / weight decay) as one fused launch, inside the timed step",
                   "l2": "no flush needed: one step streams several GB of activations (>> 126 MB L2)",
                   "step_flop_fraction_of_bf16_peak": round(step_frac, 4)},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT,
                "h2d_bytes_per_step": int((a_h.numel() + v_h.numel() + y_h.numel()) * 4), "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "arena_gb": round(m.engine().arena.bytes() / 2 ** 30, 2),
    }
    if world == 1 and not args.no_cpu_baseline:
        # bounded sample of the same workload on the host cores: ~10-20 s of CPU work (4 steps of 2 clips + 1 warm-up)
        line["cpu_baseline"] = cpu_reference_run(args.model, steps=4, warmup=1, clips=2)
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def cpu_reference_run(model_name, steps, warmup, clips):
    """The reference's algorithm on the host cores: oracle/swin_ave_oracle.py (torch CPU fp32, all threads),
    forward + soft-target CE + backward with the caller's freeze rule, on `clips` clips per step."""
    from oracle import swin_ave_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.SwinAVEConfig(img_size=224, num_frames=10, label_dim=29, **O.PRESETS[model_name])
    sd = O.make_state(cfg, seed=0)
    a, v, labels = O.make_inputs(cfg, clips, seed=0)
    for _ in range(warmup):
        O.train_step(sd, a, v, labels, cfg, "fusion")
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(sd, a, v, labels, cfg, "fusion")
    dt = time.perf_counter() - t0
    return {"value": round(clips * steps / dt, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} step(s) of {clips} clip(s) (10 frames + 10 audio images each), {model_name} fusion, fp32, "
                      f"forward + CE + backward; {dt / steps:.1f} s/step"}


def reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 4)), min(args.warmup, 1)
    cb = cpu_reference_run(args.model, steps=steps, warmup=warmup, clips=2)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": round(2 * 1e3 / cb["value"], 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD[args.model], "clips_per_step": 2, "frames": 10,
                       "note": "CPU reference arm: bounded sample of the same workload on the host cores; "
                               "steps/warmup clamped so the run ends within minutes"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="swin_t", choices=list(FLOPS_PER_CLIP))
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU (64: 401 clips/s, 76 GB arena; 32: 394; 16: ~350)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
