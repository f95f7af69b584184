#!/usr/bin/env python
"""bench.py -- adapter-train clips/sec of the AVE Swin cross-modal adapter step (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--model swin_t] [--batch B]      # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W                  # the reference's CPU path
    python bench.py --impl eager [--model swin_t] [--batch B]                       # the UNMODIFIED reference module on
                                                                                    # cuda:0, torch-eager autocast(bf16)

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
`cpu_baseline`: the UNMODIFIED reference module (staged in oracle/_ref by oracle/build_ref.py; kind "reference") timed
           on this box's host cores on a bounded sample; the oracle port (kind "port") only when oracle/_ref is absent.
`eager_baseline`: the same unmodified module on this GPU under torch-eager `autocast(bfloat16)` at the largest batch
           that fits -- the "stock B200" number of SURVEY.md 8d / BASELINE.md section 3.
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

import contextlib  # noqa: E402

import torch  # noqa: E402

METRIC = "adapter_train_clips_per_sec"
UNIT = "clips/s"
# algorithmic FLOPs per clip, forward + frozen-aware backward (BASELINE.md section 2 / SURVEY.md 8d)
FLOPS_PER_CLIP = {"swin_t": 501.6e9, "swin_b": 1587.5e9, "swin_l": 3989.0e9, "clip_b16": 1158.7e9}
WORKLOAD = {"swin_t": "AVE-29 Swin-T fusion adapter train step (BASELINE.json configs[0]/[2])",
            "swin_b": "AVE-29 Swin-B fusion adapter train step (run_adapt_ave29.py MM-Swin-AVE-Base)",
            "swin_l": "AVE-29 Swin-L fusion adapter train step (run_adapt_ave29.py MM-Swin-AVE-Large)",
            "clip_b16": "AVE-29 CLIP ViT-B/16 fusion adapter train step (BASELINE.json configs[1]; run_adapt_ave29.py "
                        "MM-CLIP-AVE-Base: 12 layers, D 768, 8 heads of 96, 197 video / 49 audio tokens per frame)"}


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
    from stgcma.config import PRESETS, CLIP_PRESETS
    from stgcma import synth
    torch.manual_seed(0)
    if name in CLIP_PRESETS:
        kw = CLIP_PRESETS[name]
        with contextlib.redirect_stdout(sys.stderr):     # the ctor prints like the reference's; stdout carries ONE json line
            m = model.CLIP_AVE.MM_CLIP_AVE(label_dim=29, num_video_frames=10, layers=kw["layers"], embed_dim=kw["embed_dim"],
                                           patch_size=kw["patch_size"], heads=kw["heads"], pretrained=None, ftmode="fusion",
                                           compute_dtype=compute_dtype)
    else:
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
    lanes, eng.lanes = getattr(eng, "lanes", 1), 1
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
    roof["traffic"], tsrc = measured_traffic(top_name.split("[")[0], B, args.model)
    roof["algorithmic_bytes_per_launch"] = round(top["bytes"] / top["calls"])
    if tsrc:
        roof["traffic_source"] = tsrc
    roof["launches"] = top["calls"]
    roof["avg_launch_ms"] = round(top["ms"] / top["calls"], 4)
    roof["time_shares"] = shares
    roof["abi_kernel_ms_per_step"] = round(tot_ms, 3)   # sum over C-ABI launches of one profiled step (serialised by the events)
    # every kernel class of the step against both roofs (GEMMs are split by contraction length: K <= 192 sits left of
    # the ridge whatever the kernel does; K >= 256 is the tensor-bound group)
    classes = {}
    for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
        sec = d["ms"] / 1e3
        classes[k.replace("stgcma_", "")] = {
            "ms": round(d["ms"], 3), "calls": d["calls"], "tflops": round(d["flops"] / sec / 1e12, 1),
            "gbs": round(d["bytes"] / sec / 1e9, 1), "tensor_frac": round(d["flops"] / sec / 1e12 / pk["tf_sustained"], 3),
            "hbm_frac": round(d["bytes"] / sec / 1e9 / pk["hbm"], 3)}
    roof["classes"] = classes
    gem = [d for k, d in prof.items() if k.startswith("stgcma_gemm[")]
    if gem:
        roof["gemm_tflops"] = round(sum(d["flops"] for d in gem) / (sum(d["ms"] for d in gem) / 1e3) / 1e12, 2)
    step_frac = value / world * FLOPS_PER_CLIP[args.model] / (pk["tf_sustained"] * 1e12)
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD[args.model], "clips_per_gpu": B, "global_batch": B * world, "frames": 10,
                   "parallelism": f"dp{world} (clips sharded, adapter-grad bucket all-reduce)" if world > 1 else "single GPU",
                   "streams": f"{getattr(m.engine(), 'lanes', 1)} micro-batch lane(s) per GPU on separate CUDA streams + dW side stream",
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
    if world == 1 and not args.no_eager_baseline:
        # the stock-B200 number: the unmodified module under torch-eager bf16 autocast on this GPU, after releasing ours
        m.engine().arena.clear()
        m.engine().frozen = m.engine().train_pack = None
        del opt, a_d, v_d, y_d, bufs
        torch.cuda.empty_cache()
        try:
            line["eager_baseline"] = eager_run(args.model, args.eager_batch or min(B, 16 if args.model.startswith("swin") else 32))
        except Exception as e:      # a reported baseline must never take the measured line down with it
            line["eager_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def reference_module(model_name):
    """The UNMODIFIED reference nn.Module (oracle/_ref, staged by oracle/build_ref.py) with the bench's synthetic weights
    and the caller's freeze rule applied; None when the staged copy is absent."""
    from oracle import shim
    from stgcma import synth
    from stgcma.config import PRESETS, CLIP_PRESETS
    if not shim.reference_available():
        return None
    torch.manual_seed(0)
    with contextlib.redirect_stdout(sys.stderr):         # the reference's ctor prints; stdout carries ONE json line
        if model_name in CLIP_PRESETS:
            kw = CLIP_PRESETS[model_name]
            ref = shim.load_reference("AVE/model/CLIP_AVE.py", "_ref_CLIP_AVE")
            m = ref.MM_CLIP_AVE(label_dim=29, layers=kw["layers"], num_video_frames=10, embed_dim=kw["embed_dim"],
                                patch_size=kw["patch_size"], heads=kw["heads"], pretrained=None, ftmode="fusion")
        else:
            ref = shim.load_reference("AVE/model/Swin_AVE.py", "_ref_Swin_AVE")
            m = ref.SwinTransformer2D_Adapter_New(label_dim=29, patch_size=[1, 4, 4], num_frames=10, window_size=7,
                                                  pretrained=None, ftmode="fusion", **PRESETS[model_name])
    synth.randomize_(m, seed=0)
    head, adapt, base = synth.split_parameters(m)
    for p in base:
        p.requires_grad_(False)
    m.train()
    return m, head, adapt


def _synth_cfg(model_name):
    from stgcma.config import PRESETS, CLIP_PRESETS, SwinAVEConfig, ClipAVEConfig
    if model_name in CLIP_PRESETS:
        return ClipAVEConfig(label_dim=29, num_frames=10, **CLIP_PRESETS[model_name])
    return SwinAVEConfig(label_dim=29, img_size=224, num_frames=10, **PRESETS[model_name])


def cpu_reference_run(model_name, steps, warmup, clips):
    """The reference on the host cores (torch CPU fp32, all threads): forward + soft-target CE + backward with the
    caller's freeze rule, on `clips` clips per step.  Runs the UNMODIFIED reference module when oracle/_ref is staged
    (kind "reference"), else the oracle port of the same algorithm (kind "port")."""
    from stgcma import synth
    torch.set_num_threads(os.cpu_count() or 1)
    built = reference_module(model_name)
    if built is not None:
        m = built[0]
        a, v, labels = synth.make_clips(_synth_cfg(model_name), clips, seed=0)
        loss_fn = torch.nn.CrossEntropyLoss()

        def one():
            m.zero_grad(set_to_none=True)
            loss_fn(m(a, v, "fusion"), labels).backward()
        kind = "reference"
    else:
        if model_name.startswith("clip"):
            from oracle import clip_ave_oracle as O
            from stgcma.config import CLIP_PRESETS
            cfg = O.ClipAVEConfig(label_dim=29, num_frames=10, **CLIP_PRESETS[model_name])
        else:
            from oracle import swin_ave_oracle as O
            cfg = O.SwinAVEConfig(img_size=224, num_frames=10, label_dim=29, **O.PRESETS[model_name])
        sd = O.make_state(cfg, seed=0)
        a, v, labels = O.make_inputs(cfg, clips, seed=0)

        def one():
            O.train_step(sd, a, v, labels, cfg, "fusion")
        kind = "port"
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return {"value": round(clips * steps / dt, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{steps} step(s) of {clips} clip(s) (10 frames + 10 audio images each), {model_name} fusion, fp32, "
                      f"forward + CE + backward; {dt / steps:.1f} s/step"}


def eager_run(model_name, batch, steps=3, warmup=2, device="cuda:0"):
    """The UNMODIFIED reference module on one B200 under torch-eager autocast(bfloat16) (the caller's loop body,
    traintest_adapt_ave29.py:158-168 with bf16 in place of fp16 -> no GradScaler needed), Adam with the caller's groups.
    The batch is halved until the step fits in memory (the reference materialises the L x L cross-modal scores)."""
    from stgcma import synth
    built = reference_module(model_name)
    if built is None:
        return {"unavailable": "oracle/_ref not staged (python -m oracle.build_ref in the build container)"}
    m, head, adapt = built
    m = m.to(device)
    opt = torch.optim.Adam([{"params": adapt, "lr": 5e-5}, {"params": head, "lr": 5e-6}], weight_decay=5e-7, betas=(0.95, 0.999))
    loss_fn = torch.nn.CrossEntropyLoss()
    cfg = _synth_cfg(model_name)
    B = batch
    while B >= 1:
        try:
            a, v, y = (t.to(device) for t in synth.make_clips(cfg, B, seed=0))

            def one():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = loss_fn(m(a, v, "fusion"), y)
                opt.zero_grad()
                loss.backward()
                opt.step()
            for _ in range(warmup):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            return {"value": round(B * steps / (ms / 1e3), 3), "unit": UNIT, "batch": B, "steps": steps, "warmup": warmup,
                    "ms_per_step": round(ms / steps, 2), "dtype": "bf16 autocast (torch eager, stock ATen / cuBLAS kernels)",
                    "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1),
                    "what": "unmodified reference module from oracle/_ref, forward + CE + backward + Adam"}
        except torch.OutOfMemoryError:
            opt.zero_grad(set_to_none=True)
            a = v = y = None
            torch.cuda.empty_cache()
            B //= 2
    return {"unavailable": "out of memory even at 1 clip"}


def eager(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = eager_run(args.model, args.eager_batch or args.batch, steps=max(1, min(args.steps, 5)), warmup=max(1, min(args.warmup, 3)))
    line = {"impl": "eager", "metric": METRIC, "unit": UNIT, "n_gpus": 1, "higher_is_better": True, "data": "synthetic",
            "config": {"workload": WORKLOAD[args.model]}}
    line.update(r)
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--model", default="swin_t", choices=list(FLOPS_PER_CLIP))
    ap.add_argument("--batch", type=int, default=64, help="clips per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true")
    ap.add_argument("--eager-batch", type=int, default=0, help="first batch the eager arm tries (halved on OOM)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference(args)
    elif args.impl == "eager":
        eager(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
