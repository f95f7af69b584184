"""The UNMODIFIED reference, executed on the GPU box beside the drop-in.

`oracle/build_ref.py` stages verbatim copies of the reference's model files, its caller
`AVE/traintest_adapt_ave29.py` and `utilities/` in the git-ignored `oracle/_ref/` (they travel with the gpurun snapshot;
/root/reference itself does not exist on the GPU box).  Two checks that no fixture can give:
  * the reference nn.Module run here on cuda in fp64 vs the drop-in (fp32 check mode 1e-3, bf16 2e-2), same weights and
    inputs -- logits and every trainable gradient;
  * the caller's real `train()` (nn.DataParallel wrap, freeze-by-name, Adam groups, fp16 autocast + GradScaler, cosine
    LR tables, validate(), checkpoint save) driving the drop-in UNCHANGED, against the same loop driving the reference
    module: train / validation losses agree, frozen weights untouched, checkpoints interchange with strict=True.
"""
import os
import re
import types

import pytest
import torch

from oracle import shim
from oracle import swin_ave_oracle as O
from oracle import clip_ave_oracle as C

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not shim.reference_available(), reason="oracle/_ref not staged (python -m oracle.build_ref)")]

SWIN_TINY = dict(img_size=112, num_frames=3, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
                 adapter_mlp_ratio=[0.25, 0.25, 0.125], label_dim=7)
CLIP_TINY = dict(label_dim=7, input_resolution=64, audio_length=480, num_frames=3, patch_size=16, embed_dim=64, layers=2, heads=4)


def _ref_swin(kw, mode):
    ref = shim.load_reference("AVE/model/Swin_AVE.py", "_ref_Swin_AVE")
    return ref.SwinTransformer2D_Adapter_New(pretrained=None, patch_size=[1, 4, 4], window_size=7, drop_path_rate=0.0, ftmode=mode, **kw)


def _ours_swin(kw, mode, dtype):
    import model
    return model.Swin_AVE.SwinTransformer2D_Adapter_New(pretrained=None, patch_size=[1, 4, 4], window_size=7, drop_path_rate=0.0,
                                                        ftmode=mode, compute_dtype=dtype, **kw)


def _clip_kw(kw):
    return dict(label_dim=kw["label_dim"], input_resolution=kw["input_resolution"], audio_length=kw["audio_length"],
                num_video_frames=kw["num_frames"], patch_size=kw["patch_size"], embed_dim=kw["embed_dim"], layers=kw["layers"],
                heads=kw["heads"], drop_path_rate=0.0, pretrained=None)


def _ref_clip(kw, mode):
    ref = shim.load_reference("AVE/model/CLIP_AVE.py", "_ref_CLIP_AVE")
    return ref.MM_CLIP_AVE(ftmode=mode, **_clip_kw(kw))


def _ours_clip(kw, mode, dtype):
    import model
    return model.CLIP_AVE.MM_CLIP_AVE(ftmode=mode, compute_dtype=dtype, **_clip_kw(kw))


FAMILIES = {
    "swin": (O, O.SwinAVEConfig, SWIN_TINY, _ref_swin, _ours_swin),
    "clip": (C, C.ClipAVEConfig, CLIP_TINY, _ref_clip, _ours_clip),
}


def rel(got, ref):
    return float((got.double().cpu() - ref.double().cpu()).abs().max() / ref.double().cpu().abs().max().clamp_min(1e-30))


def rel_l2(got, ref):
    return float((got.double().cpu() - ref.double().cpu()).norm() / ref.double().cpu().norm().clamp_min(1e-30))


def _step(m, a, v, labels, mode, is_trainable):
    m.train()
    m.mlp_head[1].eval()
    for n, p in m.named_parameters():
        p.requires_grad_(is_trainable(n))
    out = m(a, v, mode)
    loss = torch.nn.CrossEntropyLoss()(out, labels.to(out.dtype))
    loss.backward()
    torch.cuda.synchronize()
    return out.detach(), {n: p.grad.detach() for n, p in m.named_parameters() if p.requires_grad and p.grad is not None}


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_dropin_vs_unmodified_reference_module_on_gpu(family, dtype):
    Omod, Cfg, kw, make_ref, make_ours = FAMILIES[family]
    cfg = Cfg(**kw)
    sd = Omod.make_state(cfg, seed=3, bf16_exact=True)
    a, v, labels = Omod.make_inputs(cfg, 2, seed=3)
    ref = make_ref(kw, "fusion").double().cuda()
    if family == "clip":      # the reference's LayerNorm casts its input to fp32 (CLIP_AVE.py:38): its parameters must be fp32
        refmod = shim.load_reference("AVE/model/CLIP_AVE.py", "_ref_CLIP_AVE")
        for mod in ref.modules():
            if isinstance(mod, refmod.LayerNorm):
                mod.float()
    ref.load_state_dict(sd, strict=True)
    r_out, r_g = _step(ref, a.double().cuda(), v.double().cuda(), labels.double().cuda(), "fusion", Omod.is_trainable)
    ours = make_ours(kw, "fusion", dtype).cuda()
    ours.load_state_dict(sd, strict=True)
    out, g = _step(ours, a.cuda(), v.cuda(), labels.cuda(), "fusion", Omod.is_trainable)
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    metric = rel if dtype == torch.float32 else rel_l2
    assert rel(out, r_out) < tol
    assert set(r_g) <= set(g)
    gates = [n for n in r_g if "gate_" in n]
    bad = [(n, metric(g[n], r_g[n])) for n in r_g if n not in gates and not metric(g[n], r_g[n]) < tol]
    e = metric(torch.stack([g[n].reshape(()) for n in gates]), torch.stack([r_g[n].reshape(()) for n in gates]))
    assert not bad and e < tol, (sorted(bad, key=lambda t: -t[1])[:6], e)


def _caller_args(exp_dir, family):
    return types.SimpleNamespace(
        exp_dir=exp_dir, model="MM-CLIP-AVE-Base" if family == "clip" else "MM-Swin-AVE-Base", freeze_base=True, head_lr=10.0,
        lr=2e-4, lr_adapt=False, lr_cosine_adapt=True, min_lr=1e-6, n_epochs=1, warmup_epochs=0, batch_size=2, metrics="acc",
        loss="CE", dataset="ave-synthetic", n_print_steps=1, ftmode="fusion", save_model=True)


def _run_caller(family, model, sd, exp_dir, capsys):
    """One epoch (3 batches of 2 clips) of the reference's own train() + validate() on `model`."""
    Omod, Cfg, kw, _, _ = FAMILIES[family]
    cfg = Cfg(**kw)
    model.load_state_dict(sd, strict=True)
    model.mlp_head[1].p = 0.0            # head Dropout draws from torch's RNG in the reference: off on both sides
    caller = shim.load_caller()
    # The caller's GradScaler() starts at 2^16: under ITS fp16 autocast the reference module overflows on the first
    # batches and scaler.step() skips those updates, while the bf16 drop-in is finite and takes them -- the two loss
    # trajectories then differ by whole optimizer steps (measured: 3.057 vs 3.141 on the CLIP case).  Both arms therefore
    # run with the same smaller initial scale; the caller's source is untouched (only the name it imported is rebound).
    caller.GradScaler = lambda: torch.amp.GradScaler("cuda", init_scale=128.0)
    g = torch.Generator().manual_seed(5)
    a, v, _ = Omod.make_inputs(cfg, 6, seed=21)
    T, ncls = cfg.num_frames, cfg.label_dim
    hard = torch.nn.functional.one_hot(torch.randint(0, ncls, (6, T), generator=g), ncls).float()   # [B, T, classes]
    ds = torch.utils.data.TensorDataset(a, v, hard)
    loader = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False)
    os.makedirs(os.path.join(exp_dir, "models"), exist_ok=True)
    capsys.readouterr()
    caller.train(model, loader, loader, _caller_args(exp_dir, family))
    text = capsys.readouterr().out
    tl = float(re.search(r"train_loss: ([0-9.eE+-]+)", text).group(1))
    vl = float(re.search(r"valid_loss: ([0-9.eE+-]+)", text).group(1))
    return tl, vl, torch.load(os.path.join(exp_dir, "models", "audio_model.1.pth"), map_location="cpu")


@pytest.mark.parametrize("family", sorted(FAMILIES))
def test_unchanged_caller_train_loop(family, tmp_path, capsys):
    Omod, Cfg, kw, make_ref, make_ours = FAMILIES[family]
    cfg = Cfg(**kw)
    sd = Omod.make_state(cfg, seed=13)
    ours = make_ours(kw, "fusion", torch.bfloat16)
    tl, vl, ck = _run_caller(family, ours, sd, str(tmp_path / "ours"), capsys)
    ref = make_ref(kw, "fusion")
    r_tl, r_vl, r_ck = _run_caller(family, ref, sd, str(tmp_path / "ref"), capsys)
    # the reference arm runs under the caller's fp16 autocast, ours in bf16: both are ~1e-2 away from fp32
    assert abs(tl - r_tl) < 2e-2 * max(1.0, abs(r_tl)), (tl, r_tl)
    assert abs(vl - r_vl) < 2e-2 * max(1.0, abs(r_vl)), (vl, r_vl)
    # checkpoints: DataParallel prefix, same keys / shapes both ways, strict load into the other implementation
    assert all(k.startswith("module.") for k in ck) and set(ck) == set(r_ck)
    strip = lambda d: {k[len("module."):]: t for k, t in d.items()}
    make_ref(kw, "fusion").load_state_dict(strip(ck), strict=True)
    make_ours(kw, "fusion", torch.bfloat16).load_state_dict(strip(r_ck), strict=True)
    moved, frozen_same = 0, True
    for k, t in strip(ck).items():
        if not t.is_floating_point():
            continue
        if Omod.is_trainable(k):
            moved += int(not torch.equal(t, sd[k]))
        else:
            frozen_same &= torch.equal(t, sd[k])
    assert frozen_same and moved > 0.9 * sum(1 for k in sd if Omod.is_trainable(k))
    # three Adam steps move every trainable tensor by <= 3 lr-sized steps: the two runs end at nearby weights
    for k, t in strip(ck).items():
        if t.is_floating_point() and Omod.is_trainable(k) and "gate_" not in k:
            assert float((t - strip(r_ck)[k]).abs().max()) <= 3 * 2e-4 * 10.0 * 2 + 1e-6, k
