#!/usr/bin/env python
"""Multi-GPU check of the data-parallel exchange (run under torchrun on N >= 2 GPUs of one box):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_exchange_check.py

Every rank runs the tiny Swin and CLIP adapter train steps on its own clips three ways -- the one-shot peer-memory all-reduce
(`stgcma_bucket_allreduce` over torch symmetric memory), its NVLS variant (multimem.ld_reduce), the two-shot NVLS kernel
(`stgcma_bucket_allreduce_nvls`: the default above 4 ranks when the box has multicast support) and `torch.distributed.all_reduce` (NCCL) -- and the gradients must agree to fp32 summation order; they must
also equal the mean of the per-rank local gradients gathered by NCCL.  Also times the exchange alone on the Swin-T bucket size."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stg-cma_b200")]
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    import model
    from oracle import swin_ave_oracle as O
    from oracle import clip_ave_oracle as C
    from stgcma import engine as E
    ok = True
    cases = [("swin", O, dict(img_size=112, num_frames=3, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
                              adapter_mlp_ratio=[0.25, 0.25, 0.125], label_dim=7)),
             ("clip", C, dict(label_dim=7, input_resolution=64, audio_length=480, num_frames=3, patch_size=16, embed_dim=64, layers=2, heads=4))]
    for fam, mod, kw in cases:
        cfg = (O.SwinAVEConfig if fam == "swin" else C.ClipAVEConfig)(**kw)
        sd = mod.make_state(cfg, seed=3, bf16_exact=True)
        a, v, labels = mod.make_inputs(cfg, 2, seed=100 + rank)          # different clips on every rank
        grads = {}
        for mode in ("local", "nccl", "peer", "nvls1", "nvls"):
            if fam == "swin":
                m = model.Swin_AVE.SwinTransformer2D_Adapter_New(ftmode="fusion", drop_path_rate=0.0, compute_dtype=torch.float32,
                                                                 patch_size=[1, 4, 4], **kw).cuda()
            else:
                m = model.CLIP_AVE.MM_CLIP_AVE(label_dim=kw["label_dim"], input_resolution=kw["input_resolution"], audio_length=kw["audio_length"],
                                               num_video_frames=kw["num_frames"], patch_size=kw["patch_size"], embed_dim=kw["embed_dim"],
                                               layers=kw["layers"], heads=kw["heads"], drop_path_rate=0.0, pretrained=None, ftmode="fusion",
                                               compute_dtype=torch.float32).cuda()
            m.load_state_dict(sd, strict=True)
            m.train()
            m.mlp_head[1].eval()
            for n, p in m.named_parameters():
                p.requires_grad_(mod.is_trainable(n))
            m.engine().dp.mode = mode if mode != "local" else "nccl"
            m.engine().dp_allreduce = mode != "local"
            for it in range(4):        # step 3 and 4 replay CUDA graphs (2 clips <= 8): the exchange runs behind the replay
                m.zero_grad(set_to_none=True)
                out = m(a.cuda(), v.cuda(), "fusion")
                torch.nn.CrossEntropyLoss()(out, labels.cuda()).backward()
            torch.cuda.synchronize()
            grads[mode] = torch.cat([p.grad.reshape(-1) for n, p in m.named_parameters() if p.requires_grad and p.grad is not None])
            if mode == "peer":
                used = m.engine().dp.hdl is not None
            if mode == "nvls":
                nvls = m.engine().dp.hdl is not None and bool(m.engine().dp.hdl.multicast_ptr)
        mean_local = grads["local"].clone()
        dist.all_reduce(mean_local)
        mean_local /= world
        errs = {k: rel(grads[k], mean_local) for k in ("nccl", "peer", "nvls1", "nvls")}
        good = all(e < 1e-5 for e in errs.values()) and used
        ok &= good
        if rank == 0:
            print(f"{fam}: world {world}  symmetric memory used: {used}  NVLS multicast: {nvls}  max rel error vs mean of local grads: {errs}  -> {'OK' if good else 'FAIL'}", flush=True)
    # timing of the exchange alone (Swin-T bucket: 2.23 M floats)
    n = 2234000
    ex = E._DpExchange()
    for mode in ("peer", "nvls1", "nvls", "nccl"):
        ex = E._DpExchange()
        ex.mode = mode
        g = ex.bucket(n, torch.device("cuda", torch.cuda.current_device()))
        if g is None:
            g = torch.empty(n, device="cuda")
        g.normal_()
        for _ in range(3):
            ex.reduce(g)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ex.reduce(g)
        e1.record(); torch.cuda.synchronize()
        if rank == 0:
            print(f"exchange of {n * 4 / 1e6:.1f} MB, mode {mode}: {e0.elapsed_time(e1) / 20 * 1e3:.0f} us per call (incl. barriers / clone)", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
