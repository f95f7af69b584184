"""Data-parallel path on CPU: world_size 2, gloo.  Each rank runs the engine schedule (fake backend)
on half of the clips; the flat gradient bucket is all-reduced (mean) inside backward; the result must
equal the single-process gradient of the full batch."""
import os
import sys

import torch
import torch.multiprocessing as mp

TESTS = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(TESTS)
TINY = dict(img_size=112, num_frames=2, embed_dim=32, depths=[2, 2], num_heads=[1, 2], adapter_mlp_ratio=[0.25, 0.25], label_dim=5)


def _setup_paths():
    for p in (ROOT, os.path.join(ROOT, "stg-cma_b200"), TESTS):
        if p not in sys.path:
            sys.path.insert(0, p)


class _MP:
    def setattr(self, obj, name, val):
        setattr(obj, name, val)


def _run(rank, world, port, out_path):
    _setup_paths()
    import torch.distributed as dist
    import ref_ops
    import stgcma.lib
    import model
    from oracle import swin_ave_oracle as O
    torch.set_num_threads(2)
    ref_ops.install_fake_backend(_MP(), stgcma.lib)
    cfg = O.SwinAVEConfig(**TINY)
    sd = O.make_state(cfg, seed=2)
    a, v, labels = O.make_inputs(cfg, 2, seed=2)
    m = model.Swin_AVE.SwinTransformer2D_Adapter_New(ftmode="fusion", drop_path_rate=0.0, compute_dtype=torch.float64,
                                                     patch_size=[1, 4, 4], **TINY)
    m.load_state_dict(sd, strict=True)
    m.train()
    m.mlp_head[1].eval()
    for n, p in m.named_parameters():
        p.requires_grad_(O.is_trainable(n))
    T = cfg.num_frames
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        sl = slice(rank, rank + 1)
        a, v, labels = a[sl], v[sl], labels[rank * T:(rank + 1) * T]
    out = m(a, v, "fusion")
    torch.nn.CrossEntropyLoss()(out.double(), labels.double()).backward()
    g = {n: p.grad.clone() for n, p in m.named_parameters() if p.requires_grad}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        torch.save(g, out_path)


def test_two_rank_bucket_allreduce_equals_full_batch(tmp_path):
    port = 29500 + os.getpid() % 2000
    single = str(tmp_path / "single.pt")
    multi = str(tmp_path / "multi.pt")
    _run(0, 1, port, single)
    mp.spawn(_run, args=(2, port, multi), nprocs=2, join=True)
    g1, g2 = torch.load(single), torch.load(multi)
    assert set(g1) == set(g2)
    for n in g1:
        err = float((g1[n].double() - g2[n].double()).abs().max() / g1[n].double().abs().max().clamp_min(1e-30))
        assert err < 1e-5, (n, err)
