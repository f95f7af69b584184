"""CLIP-ViT variant (BASELINE.json configs[1]): the oracle restatement `oracle/clip_ave_oracle.py` replays
the fixtures produced by the UNMODIFIED reference `MM_CLIP_AVE` (oracle/make_clip_goldens.py, build container).  There is
no CUDA path for this variant yet: this pins the checker it will be built against."""
import os

import pytest
import torch

from oracle import clip_ave_oracle as C
from oracle.make_goldens import probe_index

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["clip_tiny_fusion", "clip_tiny_multimodal", "clip_b16_L2_T2_fusion"])
def test_clip_oracle_reproduces_reference(name):
    fx = torch.load(os.path.join(GOLD, name + ".pt"))
    cfg = C.ClipAVEConfig(**fx["cfg"])
    sd = C.make_state(cfg, seed=fx["seed"], dtype=torch.float64, **fx["state_kw"])
    assert list(sd.keys()) == fx["state_keys"]
    assert {k: tuple(t.shape) for k, t in sd.items()} == fx["state_shapes"]
    a, v, labels = C.make_inputs(cfg, fx["batch"], seed=fx["seed"], dtype=torch.float64)
    C.LN_FP32 = True            # the reference's LayerNorm computes in fp32 whatever the input type (CLIP_AVE.py:33-39)
    try:
        out, loss, g = C.train_step(sd, a, v, labels, cfg, fx["mode"])
    finally:
        C.LN_FP32 = False
    assert float((out - fx["logits"]).abs().max()) < 1e-10
    assert abs(float(loss - fx["loss"])) < 1e-12
    assert set(g) == set(fx["grad_norm"])
    tol = 1e-6                  # fp32 LayerNorm round-off inside an otherwise fp64 graph (oracle/validate_clip_oracle.py)
    for n, ref in fx["grads"].items():
        assert float((g[n] - ref).abs().max()) <= tol * max(1e-30, float(ref.abs().max())), n
    for n, nrm in fx["grad_norm"].items():
        assert abs(float(g[n].norm()) - nrm) <= tol * max(nrm, 1e-30), n
        probe = g[n].reshape(-1)[probe_index(g[n].numel())]
        assert float((probe - fx["grad_probe"][n]).abs().max()) <= tol * max(1e-30, float(g[n].abs().max())), n


def test_clip_geometry_and_freeze_rule():
    """ViT-B/16 as the reference instantiates it (run_adapt_ave29.py:132-139): 197 video / 49 audio tokens, r = 48;
    trainable = adapters, gates, temporal embeddings, ln_post, head."""
    cfg = C.ClipAVEConfig()
    assert (cfg.nv, cfg.na, cfg.r, cfg.audio_hw) == (197, 49, 48, (8, 6))
    names = list(C.state_shapes(cfg))
    train = [n for n in names if C.is_trainable(n)]
    assert "temporal_embedding" in train and "temporal_embedding_audio" in train and "ln_post.weight" in train
    assert "class_embedding" not in train and "conv1.weight" not in train and "ln_pre.weight" not in train
    assert not any("attn." in n or ".mlp.c_" in n or ".ln_1" in n or ".ln_2" in n for n in train)
    n_adapter = sum(torch.Size(s).numel() for n, s in C.state_shapes(cfg).items() if "Adapter" in n or "gate_" in n)
    assert abs(n_adapter / 1e6 - 5.38) < 0.05      # 5.38 M adapter parameters


def test_clip_fusion_with_zero_gates_is_multimodal():
    cfg = C.ClipAVEConfig(label_dim=7, input_resolution=64, audio_length=480, num_frames=3, embed_dim=64, layers=2, heads=4)
    sd = C.make_state(cfg, seed=1, dtype=torch.float64)
    for k in sd:
        if "gate_" in k:
            sd[k] = torch.zeros_like(sd[k])
    a, v, _ = C.make_inputs(cfg, 2, seed=1, dtype=torch.float64)
    assert torch.equal(C.forward(sd, a, v, cfg, "fusion"), C.forward(sd, a, v, cfg, "multimodal"))
