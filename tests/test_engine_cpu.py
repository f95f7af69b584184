"""Host-side schedule (stgcma/engine.py + model/Swin_AVE.py) against the oracle, on CPU.

The ctypes kernel wrappers are replaced by the per-op torch references of tests/ref_ops.py, so
this checks everything ABOVE the C ABI: the order of operations, residual wiring, what is saved
for backward, grouped per-modality adapters, the flat gradient bucket and its offsets, drop-path
and head-dropout plumbing.  The kernels themselves are checked op by op in the `-m gpu` tests.
"""
import pytest
import torch

import ref_ops
from oracle import swin_ave_oracle as O

TINY = dict(img_size=112, num_frames=3, embed_dim=32, depths=[2, 2, 2], num_heads=[1, 2, 4],
            adapter_mlp_ratio=[0.25, 0.25, 0.125], label_dim=7)


def build(cfg_kw, mode, dtype, drop_path_rate=0.0):
    import model
    import stgcma.lib
    m = model.Swin_AVE.SwinTransformer2D_Adapter_New(
        label_dim=cfg_kw["label_dim"], img_size=cfg_kw["img_size"], num_frames=cfg_kw["num_frames"],
        embed_dim=cfg_kw["embed_dim"], depths=cfg_kw["depths"], num_heads=cfg_kw["num_heads"],
        adapter_mlp_ratio=cfg_kw["adapter_mlp_ratio"], ftmode=mode, drop_path_rate=drop_path_rate,
        compute_dtype=dtype, patch_size=[1, 4, 4], window_size=7, pretrained=None)
    return m


def run_model(m, sd, a, v, labels, mode):
    m.load_state_dict(sd, strict=True)
    for n, p in m.named_parameters():
        p.requires_grad_(O.is_trainable(n))
    out = m(a, v, mode)
    loss = torch.nn.CrossEntropyLoss()(out.double(), labels.double())
    loss.backward()
    return out.detach(), loss.detach(), {n: p.grad for n, p in m.named_parameters() if p.requires_grad}


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("fused_mlp", [False, True])
@pytest.mark.parametrize("mode", ["fusion", "multimodal"])
def test_schedule_matches_oracle(monkeypatch, mode, fused_mlp):
    """fused_mlp: the bf16 production schedule (stgcma_mlp_fwd / _bwd instead of layernorm + 3 gemm calls per
    direction), exercised here in fp64 through the torch references of those two entry points."""
    import stgcma.lib
    ref_ops.install_fake_backend(monkeypatch, stgcma.lib)
    monkeypatch.setattr(stgcma.lib, "mlp_fused_ok", lambda *a, **k: fused_mlp)
    cfg = O.SwinAVEConfig(**TINY)
    sd = O.make_state(cfg, seed=3, dtype=torch.float64)
    a, v, labels = O.make_inputs(cfg, 2, seed=3, dtype=torch.float64)
    o_out, o_loss, o_g = O.train_step(sd, a, v, labels, cfg, mode)
    m = build(TINY, mode, torch.float64)
    m.train()
    m.mlp_head[1].eval()
    sd32 = {k: (t.float() if t.is_floating_point() else t) for k, t in sd.items()}
    out, loss, g = run_model(m, sd32, a.float(), v.float(), labels, mode)
    assert out.dtype == torch.float32 and tuple(out.shape) == (6, 7)
    assert rel(out, o_out) < 2e-6
    trainable = {n for n in sd if O.is_trainable(n)}
    if mode == "multimodal":  # gates exist but are unused -> no gradient in the reference either
        trainable = {n for n in trainable if "gate_" not in n}
    assert trainable <= set(g)
    for n in sorted(trainable):
        assert g[n] is not None, n
        assert rel(g[n], o_g[n]) < 2e-5, n


def test_schedule_matches_reference_golden(monkeypatch):
    """Same check against the fixture produced by the UNMODIFIED reference (oracle/make_goldens.py)."""
    import os
    import stgcma.lib
    ref_ops.install_fake_backend(monkeypatch, stgcma.lib)
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "tiny3stage_fusion.pt"))
    cfg = O.SwinAVEConfig(**fx["cfg"])
    sd = O.make_state(cfg, seed=fx["seed"], dtype=torch.float32, **fx["state_kw"])
    a, v, labels = O.make_inputs(cfg, fx["batch"], seed=fx["seed"])
    m = build(fx["cfg"], fx["mode"], torch.float64)
    m.train()
    m.mlp_head[1].eval()
    out, loss, g = run_model(m, sd, a, v, labels, fx["mode"])
    assert rel(out, fx["logits"]) < 2e-6
    for n, ref in fx["grads"].items():
        assert rel(g[n], ref) < 2e-5, n


def test_drop_path_and_head_dropout(monkeypatch):
    """Stochastic layers: the engine draws the masks, the oracle replays them."""
    import stgcma.lib
    ref_ops.install_fake_backend(monkeypatch, stgcma.lib)
    cfg = O.SwinAVEConfig(**TINY)
    sd = O.make_state(cfg, seed=5, dtype=torch.float32)
    a, v, labels = O.make_inputs(cfg, 2, seed=5)
    m = build(TINY, "fusion", torch.float64, drop_path_rate=0.5)
    m.train()
    m.load_state_dict(sd, strict=True)
    for n, p in m.named_parameters():
        p.requires_grad_(O.is_trainable(n))
    torch.manual_seed(0)
    out = m(a, v, "fusion")
    saved = m.engine().saved_masks(out.grad_fn.saved) if hasattr(out.grad_fn, "saved") else None
    assert saved is not None
    B, T = 2, cfg.num_frames
    masks = {}
    for (s, i), dp in saved["dp"].items():
        if dp is not None:
            L = cfg.stage_res(s) ** 2
            masks[(s, i)] = dp.view(2, B, T, L)[:, :, 0, :].reshape(2, B * L)
    assert masks, "drop_path_rate=0.5 must activate at least one mask"
    assert any(float((mk == 0).float().mean()) > 0.1 for mk in masks.values())
    keep = saved["keepmask"]
    assert keep is not None and set(keep.unique().tolist()) <= {0.0, 2.0}
    full_masks = {k: masks.get(k, torch.ones(2, B * cfg.stage_res(k[0]) ** 2)) for k in
                  [(s, i) for s in range(cfg.num_stages) for i in range(cfg.depths[s]) if i % 2 == 0]}
    sd64 = {k: (t.double() if t.is_floating_point() else t) for k, t in sd.items()}
    loss = torch.nn.CrossEntropyLoss()(out.double(), labels.double())
    loss.backward()
    o_out, o_loss, o_g = O.train_step(sd64, a.double(), v.double(), labels.double(), cfg, "fusion",
                                      drop_masks={k: mk.double() for k, mk in full_masks.items()},
                                      head_keep=keep.double())
    assert rel(out, o_out) < 2e-6
    for n, p in m.named_parameters():
        if p.requires_grad:
            assert rel(p.grad, o_g[n]) < 2e-5, n


def test_refuses_unfrozen_backbone_and_cpu(monkeypatch):
    import stgcma.lib
    m = build(TINY, "fusion", torch.float32)
    cfg = O.SwinAVEConfig(**TINY)
    a, v, _ = O.make_inputs(cfg, 1, seed=0)
    with pytest.raises(stgcma.lib.StgcmaError):       # CPU tensors: no fallback
        m(a, v, "fusion")
    ref_ops.install_fake_backend(monkeypatch, stgcma.lib)
    with pytest.raises(stgcma.lib.StgcmaError):       # backbone not frozen
        m(a, v, "fusion")
    with pytest.raises(NotImplementedError):
        build(TINY, "videoonly", torch.float32)


def test_pretrained_checkpoint_route_matches_reference(tmp_path, capsys):
    """`pretrained=<path>` (the route every real run takes, run_adapt_ave29.py:156-165): a synthetic ImageNet-style Swin
    checkpoint -- 2-D patch conv, no audio stream, no adapters, one unexpected key -- loaded by the drop-in and by the
    UNMODIFIED reference (Swin_AVE.py:1363-1413) must give identical state dicts: conv inflation over time, audio conv =
    RGB mean, audio patch-norm copied, adapter D_fc2 / gates zero afterwards, same Missing / Unexpected key report."""
    from oracle import shim
    if not shim.reference_available():
        pytest.skip("the unmodified reference is neither at /root/reference nor staged in oracle/_ref")
    ref = shim.load_reference()
    kw = dict(label_dim=7, patch_size=[1, 4, 4], num_frames=3, embed_dim=32, depths=[2, 2], num_heads=[1, 2], window_size=7,
              ftmode="fusion", adapter_mlp_ratio=[0.25, 0.25], img_size=112)
    torch.manual_seed(0)
    donor = ref.SwinTransformer2D_Adapter_New(pretrained=None, **kw)
    g = torch.Generator().manual_seed(1)
    ck = {}
    for k, t in donor.state_dict().items():
        if any(s in k for s in ("Adapter", "gate_", "temporal_position", "t_relative", "_audio", "mlp_head")):
            continue   # an ImageNet Swin checkpoint has none of these
        ck[k] = torch.randn(t.shape, generator=g) * 0.05 if t.is_floating_point() and "attn_mask" not in k else t.clone()
    ck["patch_embed.proj.weight"] = torch.randn(32, 3, 4, 4, generator=g) * 0.05   # 2-D conv, inflated by the loader
    ck["head.weight"] = torch.randn(5, 64, generator=g)                            # unexpected key: reported, ignored
    path = str(tmp_path / "swin_synth.pth")
    torch.save({"model": ck}, path)

    torch.manual_seed(2)
    r = ref.SwinTransformer2D_Adapter_New(pretrained=path, **kw)
    ref_log = capsys.readouterr().out
    import model
    torch.manual_seed(2)
    m = model.Swin_AVE.SwinTransformer2D_Adapter_New(pretrained=path, **kw)
    our_log = capsys.readouterr().out
    rs, ms = r.state_dict(), m.state_dict()
    assert list(rs.keys()) == list(ms.keys())
    loaded = set(ck) | {"patch_embed_audio.proj.weight", "patch_embed_audio.proj.bias", "patch_embed_audio.norm.weight", "patch_embed_audio.norm.bias"}
    for k in rs:
        if k in loaded or "D_fc2" in k or "gate_" in k or not rs[k].is_floating_point():
            assert torch.equal(rs[k], ms[k]), k   # checkpoint-derived tensors and the zero-initialised ones: bit equal
        else:
            assert rs[k].shape == ms[k].shape, k  # randomly initialised (trunc-normal draws differ in order)
    assert ms["patch_embed.proj.weight"].shape == (32, 3, 1, 4, 4) and ms["patch_embed_audio.proj.weight"].shape == (32, 1, 1, 4, 4)
    assert torch.equal(ms["patch_embed_audio.proj.weight"], ms["patch_embed.proj.weight"].mean(dim=1, keepdim=True))

    def report(log, tag):
        line = next(l for l in log.splitlines() if l.startswith(tag))
        return sorted(eval(line[len(tag):]))
    for tag in ("Missing keys: ", "Unexpected keys: "):
        assert report(ref_log, tag) == report(our_log, tag), tag
    assert report(our_log, "Unexpected keys: ") == ["head.weight"]
