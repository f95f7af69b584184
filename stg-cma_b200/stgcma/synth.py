"""Seeded synthetic weights and clips for benchmarks / smoke runs (no datasets or checkpoints are
reachable offline).  Shapes follow SURVEY.md section 8d: video ~ N(0,1) [B,3,T,224,224], Swin audio
~ N(0,1) [B,T,224,224], soft labels = softmax(N(0,1)) [B*T, classes]."""
import torch

_ADAPT_KEYS = ("adapter", "temporal_embedding", "ln_post", "Adapter", "my_tokens", "gate_", "ln_before",
               "temporal_position_bias_table")
_HEAD = {"mlp_head.0.weight", "mlp_head.0.bias", "mlp_head.2.weight", "mlp_head.2.bias"}


def split_parameters(model):
    """The caller's grouping-by-name (traintest_adapt_ave29.py:38-55): (head, adapters, frozen base)."""
    head, adapt, base = [], [], []
    for name, p in model.named_parameters():
        if name in _HEAD:
            head.append(p)
        elif any(k in name for k in _ADAPT_KEYS):
            adapt.append(p)
        else:
            base.append(p)
    return head, adapt, base


def randomize_(model, seed=0):
    """The reference zero-initialises every adapter D_fc2 and the gates, which makes the whole
    cross-modal path and dW(D_fc1) identically zero; benchmarks use live values instead."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if "D_fc2.weight" in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
            elif "gate_" in name:
                p.copy_(0.25 + 0.5 * torch.rand(p.shape, generator=g))
            elif ("norm" in name or ".ln_" in name or name.startswith("ln_")) and name.endswith("weight"):
                p.copy_(1.0 + 0.05 * torch.randn(p.shape, generator=g))
    return model


def make_clips(cfg, batch, seed=0, pin=False):
    g = torch.Generator().manual_seed(1000 + seed)
    T = cfg.num_frames
    if hasattr(cfg, "audio_shape"):      # CLIP path: audio is the [audio_length / 10, 128] fbank slice per frame
        S = cfg.input_resolution
        v = torch.randn(batch, 3, T, S, S, generator=g)
        a = torch.randn(batch, T, *cfg.audio_shape, generator=g)
    else:
        S = cfg.img_size
        v = torch.randn(batch, 3, T, S, S, generator=g)
        a = torch.randn(batch, T, S, S, generator=g)
    labels = torch.softmax(torch.randn(batch * T, cfg.label_dim, generator=g), -1)
    if pin:
        v, a, labels = v.pin_memory(), a.pin_memory(), labels.pin_memory()
    return a, v, labels
