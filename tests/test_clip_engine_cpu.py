"""Host-side schedule of the CLIP-ViT variant (stgcma/clip_engine.py + model/CLIP_AVE.py) against the oracle, on CPU.

As in test_engine_cpu.py the ctypes kernel wrappers are replaced by the per-op torch references of tests/ref_ops.py:
this checks everything ABOVE the C ABI -- operation order, residual wiring, the per-modality row slices (197-token video
block / 49-token audio block), what is saved for backward, the gradient bucket, DropPath / dropout plumbing -- against
`oracle/clip_ave_oracle.py` and against the fixtures produced by the UNMODIFIED reference `MM_CLIP_AVE`.
"""
import os

import pytest
import torch

import ref_ops
from oracle import clip_ave_oracle as C

TINY = dict(label_dim=7, input_resolution=64, audio_length=480, num_frames=3, patch_size=16, embed_dim=64, layers=2, heads=4)
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def build(kw, mode, dtype, drop_path_rate=0.0):
    import model
    return model.CLIP_AVE.MM_CLIP_AVE(label_dim=kw["label_dim"], input_resolution=kw["input_resolution"],
                                      audio_length=kw["audio_length"], num_video_frames=kw["num_frames"],
                                      patch_size=kw["patch_size"], embed_dim=kw["embed_dim"], layers=kw["layers"],
                                      heads=kw["heads"], drop_path_rate=drop_path_rate, pretrained=None, ftmode=mode,
                                      compute_dtype=dtype)


def run_model(m, sd, a, v, labels, mode):
    m.load_state_dict(sd, strict=True)
    for n, p in m.named_parameters():
        p.requires_grad_(C.is_trainable(n))
    out = m(a, v, mode)
    loss = torch.nn.CrossEntropyLoss()(out.double(), labels.double())
    loss.backward()
    return out.detach(), loss.detach(), {n: p.grad for n, p in m.named_parameters() if p.requires_grad}


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def test_state_dict_names_match_reference():
    """Names, shapes and ORDER of the state_dict are the reference's (oracle.state_shapes is checked against the
    unmodified reference by oracle/validate_clip_oracle.py and stored in the fixtures)."""
    fx = torch.load(os.path.join(GOLD, "clip_b16_L2_T2_fusion.pt"))
    m = build(fx["cfg"], "fusion", torch.float32)
    sd = m.state_dict()
    assert list(sd.keys()) == fx["state_keys"]
    assert {k: tuple(t.shape) for k, t in sd.items()} == fx["state_shapes"]
    names = {n for n, _ in m.engine().trainable_params()}
    assert names == {n for n in sd if C.is_trainable(n)}


@pytest.mark.parametrize("mode", ["fusion", "multimodal"])
def test_schedule_matches_oracle(monkeypatch, mode):
    import stgcma.lib
    ref_ops.install_fake_backend(monkeypatch, stgcma.lib)
    cfg = C.ClipAVEConfig(**TINY)
    sd = C.make_state(cfg, seed=3, dtype=torch.float64)
    a, v, labels = C.make_inputs(cfg, 2, seed=3, dtype=torch.float64)
    o_out, o_loss, o_g = C.train_step(sd, a, v, labels, cfg, mode)
    m = build(TINY, mode, torch.float64)
    m.train()
    m.mlp_head[1].eval()
    sd32 = {k: t.float() for k, t in sd.items()}
    out, loss, g = run_model(m, sd32, a.float(), v.float(), labels, mode)
    assert out.dtype == torch.float32 and tuple(out.shape) == (6, 7)
    assert rel(out, o_out) < 2e-6
    trainable = {n for n in sd if C.is_trainable(n)}
    if mode == "multimodal":
        trainable = {n for n in trainable if "gate_" not in n}
    assert trainable <= set(g)
    for n in sorted(trainable):
        assert g[n] is not None, n
        assert rel(g[n], o_g[n]) < 2e-5, n


@pytest.mark.parametrize("name", ["clip_tiny_fusion", "clip_tiny_multimodal"])
def test_schedule_matches_reference_golden(monkeypatch, name):
    import stgcma.lib
    ref_ops.install_fake_backend(monkeypatch, stgcma.lib)
    fx = torch.load(os.path.join(GOLD, name + ".pt"))
    cfg = C.ClipAVEConfig(**fx["cfg"])
    sd = C.make_state(cfg, seed=fx["seed"], dtype=torch.float32, **fx["state_kw"])
    a, v, labels = C.make_inputs(cfg, fx["batch"], seed=fx["seed"])
    m = build(fx["cfg"], fx["mode"], torch.float64)
    m.train()
    m.mlp_head[1].eval()
    out, loss, g = run_model(m, sd, a, v, labels, fx["mode"])
    assert rel(out, fx["logits"]) < 2e-6
    for n, ref in fx["grads"].items():
        assert rel(g[n], ref) < 2e-5, n


def test_drop_path_and_head_dropout(monkeypatch):
    """Stochastic layers: the engine draws the masks (DropPath per TOKEN INDEX, as the reference's DropPath on the
    sequence-first tensor does, CLIP_AVE.py:372,377), the oracle replays them."""
    import stgcma.lib
    ref_ops.install_fake_backend(monkeypatch, stgcma.lib)
    cfg = C.ClipAVEConfig(**TINY)
    sd = C.make_state(cfg, seed=5, dtype=torch.float32)
    a, v, labels = C.make_inputs(cfg, 2, seed=5)
    m = build(TINY, "fusion", torch.float64, drop_path_rate=0.5)
    m.train()
    m.load_state_dict(sd, strict=True)
    for n, p in m.named_parameters():
        p.requires_grad_(C.is_trainable(n))
    torch.manual_seed(0)
    out = m(a, v, "fusion")
    saved = m.engine().saved_masks(out.grad_fn.saved)
    assert saved["dp"], "drop_path_rate=0.5 must activate at least one mask"
    B = 2
    masks = {i: dict(v=d["v"].repeat(B).double(), a=d["a"].repeat(B).double()) for i, d in saved["dp"].items()}   # (b n) order
    assert any(float((d["v"] == 0).float().mean()) > 0.1 for d in saved["dp"].values())
    keep = saved["keepmask"]
    assert keep is not None and set(keep.unique().tolist()) <= {0.0, 2.0}
    loss = torch.nn.CrossEntropyLoss()(out.double(), labels.double())
    loss.backward()
    sd64 = {k: t.double() for k, t in sd.items()}
    o_out, _, o_g = C.train_step(sd64, a.double(), v.double(), labels.double(), cfg, "fusion", drop_masks=masks,
                                 head_keep=keep.double())
    assert rel(out, o_out) < 2e-6
    for n, p in m.named_parameters():
        if p.requires_grad:
            assert rel(p.grad, o_g[n]) < 2e-5, n


def test_refuses_unfrozen_backbone_and_cpu(monkeypatch):
    import stgcma.lib
    m = build(TINY, "fusion", torch.float32)
    cfg = C.ClipAVEConfig(**TINY)
    a, v, _ = C.make_inputs(cfg, 1, seed=0)
    with pytest.raises(stgcma.lib.StgcmaError):       # CPU tensors: no fallback
        m(a, v, "fusion")
    ref_ops.install_fake_backend(monkeypatch, stgcma.lib)
    with pytest.raises(stgcma.lib.StgcmaError):       # backbone not frozen
        m(a, v, "fusion")
    with pytest.raises(NotImplementedError):
        build(TINY, "videoonly", torch.float32)
