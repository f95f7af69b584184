#!/usr/bin/env python
"""DRAM traffic per C-ABI kernel class from an ncu launch list that carries dram__bytes_read.sum / dram__bytes_write.sum
(and gpu__time_duration.sum) for every launch of one train step:

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/traffic.csv python scripts/profile_step.py --batch 32
    python scripts/traffic_summary.py gpurun_out/traffic.csv 32 > profiles/rNN_traffic_swin_t_b32.json

bench.py reads the newest profiles/*_traffic_*.json whose batch matches and reports `roofline.traffic` = measured DRAM
bytes per launch of the dominant class (same per-launch basis as `achieved`)."""
import csv
import json
import re
import sys

CLASS = [  # kernel-name regex -> C-ABI entry point (the key bench.py's live profile uses)
    (r"gemm_tc_kernel|skinny_kernel|gemm_simt_kernel", "stgcma_gemm"),
    (r"gemm_tn_kernel|gemm_tn_simt_kernel", "stgcma_gemm_tn"),
    (r"xg_bwd_kernel|xmodal_prep_kernel|xmodal_global_bwd", "stgcma_xmodal_global_bwd"),
    (r"xg_fwd_kernel|xmodal_global_fwd", "stgcma_xmodal_global_fwd"),
    (r"window_bwd_kernel", "stgcma_window_attn_bwd"), (r"window_fwd_kernel", "stgcma_window_attn_fwd"),
    (r"ln_bwd", "stgcma_layernorm_bwd"), (r"ln_fwd", "stgcma_layernorm_fwd"),
    (r"mlp_bwd_kernel", "stgcma_mlp_bwd"), (r"mlp_fwd_kernel", "stgcma_mlp_fwd"),
    (r"tt_bwd_kernel", "stgcma_temporal_attn_bwd"), (r"tt_fwd_kernel", "stgcma_temporal_attn_fwd"),
    (r"xw_bwd_kernel", "stgcma_xmodal_window_bwd"), (r"xw_fwd_kernel", "stgcma_xmodal_window_fwd"),
    (r"patch_embed", "stgcma_patch_embed_fwd"), (r"pack_params", "stgcma_pack_params"),
    (r"head_pool_bwd", "stgcma_head_pool_bwd"), (r"head_pool_fwd", "stgcma_head_pool_fwd"),
    # CLIP-ViT variant: the temporal MHA runs the tt_* tile kernels (classified above as temporal_attn when the Swin
    # model runs; `main` re-labels them for a clip workload)
    (r"mha_tc_bwd|mha_bwd_dq|mha_bwd_dkv|mha_bwd_simt", "stgcma_mha_bwd"), (r"mha_tc_fwd|mha_fwd_kernel|mha_fwd_simt", "stgcma_mha_fwd"),
    (r"im2col", "stgcma_im2col"), (r"clip_embed", "stgcma_clip_embed_fwd"), (r"clip_temb", "stgcma_clip_temb_bwd"),
    (r"cls_ln_bwd", "stgcma_cls_ln_bwd"), (r"cls_ln_fwd", "stgcma_cls_ln_fwd"), (r"adam_kernel", "stgcma_adam_step"),
]


def main(path, batch, workload="swin_t"):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, im, iu, iv, iid = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}   # ncu: decimal units
    per = {}
    for r in rd:
        try:
            v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1.0)
        except (ValueError, IndexError):
            continue
        d = per.setdefault(r[iid], dict(name=r[ik], bytes=0.0, ns=0.0))
        if r[im].startswith("dram__bytes"):
            d["bytes"] += v
        elif r[im].startswith("gpu__time_duration"):
            d["ns"] += v
    out = {}
    for d in per.values():
        cls = next((c for rx, c in CLASS if re.search(rx, d["name"])), "other")
        if workload.startswith("clip") and cls.startswith("stgcma_temporal_attn"):
            cls = cls.replace("temporal_attn", "mha")
        o = out.setdefault(cls, dict(launches=0, dram_bytes=0.0, ms=0.0))
        o["launches"] += 1
        o["dram_bytes"] += d["bytes"]
        o["ms"] += d["ns"] / 1e6
    for o in out.values():
        o["dram_bytes_per_launch"] = round(o["dram_bytes"] / o["launches"])
        o["dram_bytes"] = round(o["dram_bytes"])
        o["ms"] = round(o["ms"], 3)
    print(json.dumps({"workload": f"{workload} fusion train step", "clips_per_gpu": batch,
                      "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, one profiled step "
                                "(--clock-control none; launches serialised, caches cold between replays)",
                      "classes": dict(sorted(out.items(), key=lambda kv: -kv[1]["ms"]))}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), *(sys.argv[3:4]))
