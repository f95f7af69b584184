"""TEST INFRASTRUCTURE ONLY -- parity oracle for the CLIP-ViT variant of the AVE adapter train step
(BASELINE.json configs[1]; SURVEY.md section 8a, row "CLIP: ResidualAttentionBlock.forward fusion").

A CPU restatement (plain torch, fp32 / fp64, autograd for the gradients) of

    MM_CLIP_AVE.forward(a, v, 'fusion' | 'multimodal')          AVE/model/CLIP_AVE.py:1089-1140 / :1038-1087
      patch embed (conv1 / conv1_audio, no bias), class token,
      positional + temporal embeddings, ln_pre                  :1090-1120
      ResidualAttentionBlock.forward, mode 'fusion_adapt'       :359-429   ('multimodal_adapt_no_fusion' :256-289)
      ResidualAttentionBlock.attention (nn.MultiheadAttention)  :106-108
      Adapter                                                    :13-31
      head: ln_post, class token, cat(a, v), mlp_head            :1124-1139

written functionally over a flat ``state`` dict that uses the reference's own state_dict names.  The product
(`stg-cma_b200/`) has NO CLIP path yet: this file and its fixtures are the first step of that
row -- the checker the kernels will be built against.  Only tests may import it.

PARITY PINNING: like the Swin oracle, pinned against the unmodified reference itself by
`python -m oracle.validate_clip_oracle` (build container only, through oracle/shim.py); reference outputs are stored
as fixtures by `oracle/make_clip_goldens.py` and replayed by `tests/test_clip_oracle_golden.py` anywhere.
"""
from collections import OrderedDict
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class ClipAVEConfig:
    label_dim: int = 29
    input_resolution: int = 224
    audio_length: int = 1024        # target_length of the fbank; the model sees [B, T, audio_length // 10, 128]
    num_frames: int = 10
    patch_size: int = 16
    embed_dim: int = 768
    layers: int = 12
    heads: int = 8                  # the reference instantiates ViT-B/16 with 8 heads (run_adapt_ave29.py:137)
    adapter_ratio: float = 0.0625   # Adapter(mlp_ratio=0.0625), CLIP_AVE.py:14

    @property
    def nv(self):                   # video tokens per frame incl. the class token
        return (self.input_resolution // self.patch_size) ** 2 + 1

    @property
    def audio_hw(self):             # (f_dim, t_dim) of get_shape_a (:793-800): conv 16x16 stride patch on [128, audio_length/10]
        tdim = int(self.audio_length * (1 / 10))
        return ((128 - 16) // self.patch_size + 1, (tdim - 16) // self.patch_size + 1)

    @property
    def na(self):
        f, t = self.audio_hw
        return f * t + 1

    @property
    def r(self):
        return int(self.embed_dim * self.adapter_ratio)


_ADAPT_KEYS = ("adapter", "temporal_embedding", "ln_post", "Adapter", "my_tokens", "gate_", "ln_before",
               "temporal_position_bias_table")
_MLP_LIST = ("mlp_head.0.weight", "mlp_head.0.bias", "mlp_head.2.weight", "mlp_head.2.bias")


def is_trainable(name):
    """The caller's freeze rule, traintest_adapt_ave29.py:38-61 (freeze_base=True)."""
    return name in _MLP_LIST or any(k in name for k in _ADAPT_KEYS)


def state_shapes(cfg: ClipAVEConfig):
    """Reference state_dict names -> shapes for ftmode 'fusion' / 'multimodal' (identical modules in both)."""
    D, r, P = cfg.embed_dim, cfg.r, cfg.patch_size
    sh = OrderedDict()
    sh["class_embedding"] = (D,)
    sh["positional_embedding"] = (cfg.nv, D)
    sh["positional_embedding_audio"] = (cfg.na, D)
    sh["temporal_embedding"] = (1, cfg.num_frames, D)
    sh["temporal_embedding_audio"] = (1, cfg.num_frames, D)
    sh["conv1.weight"] = (D, 3, P, P)
    sh["ln_pre.weight"], sh["ln_pre.bias"] = (D,), (D,)
    sh["conv1_audio.weight"] = (D, 1, P, P)
    for i in range(cfg.layers):
        p = f"transformer.resblocks.{i}."
        sh[p + "gate_v"], sh[p + "gate_a"] = (1,), (1,)
        sh[p + "attn.in_proj_weight"], sh[p + "attn.in_proj_bias"] = (3 * D, D), (3 * D,)
        sh[p + "attn.out_proj.weight"], sh[p + "attn.out_proj.bias"] = (D, D), (D,)
        sh[p + "ln_1.weight"], sh[p + "ln_1.bias"] = (D,), (D,)
        sh[p + "mlp.c_fc.weight"], sh[p + "mlp.c_fc.bias"] = (4 * D, D), (4 * D,)
        sh[p + "mlp.c_proj.weight"], sh[p + "mlp.c_proj.bias"] = (D, 4 * D), (D,)
        sh[p + "ln_2.weight"], sh[p + "ln_2.bias"] = (D,), (D,)
        for ad in ("MLP_Adapter", "S_Adapter", "T_Adapter", "S_Adapter_Audio", "MLP_Adapter_Audio", "T_Adapter_Audio"):
            sh[p + ad + ".D_fc1.weight"], sh[p + ad + ".D_fc1.bias"] = (r, D), (r,)
            sh[p + ad + ".D_fc2.weight"], sh[p + ad + ".D_fc2.bias"] = (D, r), (D,)
    sh["ln_post.weight"], sh["ln_post.bias"] = (D,), (D,)
    sh["mlp_head.0.weight"], sh["mlp_head.0.bias"] = (512, 2 * D), (512,)
    sh["mlp_head.2.weight"], sh["mlp_head.2.bias"] = (cfg.label_dim, 512), (cfg.label_dim,)
    return sh


def make_state(cfg: ClipAVEConfig, seed=0, dtype=torch.float32, bf16_exact=False):
    """Seeded synthetic weights: fan-in scaled normals, LN gamma ~ 1, adapters' D_fc2 NON-zero and gates
    in [0.25, 0.75] (the reference's zero init makes the cross-modal path and half the gradients vanish at step 0)."""
    g = torch.Generator().manual_seed(7000 + seed)
    sd = OrderedDict()
    for name, shape in state_shapes(cfg).items():
        if "gate_" in name:
            t = 0.25 + 0.5 * torch.rand(shape, generator=g, dtype=torch.float64)
        elif name.startswith("ln_") and name.endswith("weight") or ".ln_" in name and name.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif name.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif "embedding" in name:
            t = 0.05 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif "conv1" in name:
            fan = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g, dtype=torch.float64) * fan ** -0.5
        elif "D_fc2.weight" in name:
            t = 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            t = torch.randn(shape, generator=g, dtype=torch.float64) * shape[-1] ** -0.5
        if bf16_exact:
            t = t.to(torch.bfloat16).to(torch.float64)
        sd[name] = t.to(dtype)
    return sd


def make_inputs(cfg: ClipAVEConfig, batch, seed=0, dtype=torch.float32):
    """v ~ N(0,1) [B,3,T,H,W]; a ~ N(0,1) [B,T,audio_length/10,128] (AVE/dataloader.py:244-256); soft labels [B*T, C]."""
    g = torch.Generator().manual_seed(9000 + seed)
    R = cfg.input_resolution
    v = torch.randn(batch, 3, cfg.num_frames, R, R, generator=g, dtype=torch.float64).to(dtype)
    a = torch.randn(batch, cfg.num_frames, int(cfg.audio_length * (1 / 10)), 128, generator=g, dtype=torch.float64).to(dtype)
    labels = torch.softmax(torch.randn(batch * cfg.num_frames, cfg.label_dim, generator=g, dtype=torch.float64), -1).to(dtype)
    return a, v, labels


# ---------------------------------------------------------------------------------------------------------------
LN_FP32 = False   # the reference's LayerNorm subclass (:33-39) always computes in fp32 and casts back; set by the fp64
                  # validation so that both sides round identically (irrelevant when the oracle itself runs in fp32)


def _layer_norm(x, w, b):
    if LN_FP32 and x.dtype != torch.float32:
        return F.layer_norm(x.float(), (x.shape[-1],), w.float(), b.float(), 1e-5).to(x.dtype)
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _ln(x, sd, p):
    return _layer_norm(x, sd[p + ".weight"], sd[p + ".bias"])


def _mha(x, sd, p, heads):
    """nn.MultiheadAttention(d, heads)(x, x, x) on a sequence-first [N, B, D] tensor, no mask (:106-108)."""
    N, Bn, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)

    def split(t):   # [N, B, D] -> [B * heads, N, hd]
        return t.reshape(N, Bn * heads, hd).transpose(0, 1)
    q, k, v = split(q) * hd ** -0.5, split(k), split(v)
    att = torch.softmax(q @ k.transpose(1, 2), dim=-1)
    o = (att @ v).transpose(0, 1).reshape(N, Bn, D)
    return F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])


def _adapter_hidden(x, sd, p):
    return F.gelu(F.linear(x, sd[p + ".D_fc1.weight"], sd[p + ".D_fc1.bias"]))


def _adapter_out(h, sd, p):
    return F.linear(h, sd[p + ".D_fc2.weight"], sd[p + ".D_fc2.bias"])


def _mlp(x, sd, p):
    h = F.linear(x, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])
    h = h * torch.sigmoid(1.702 * h)                                   # QuickGELU (:41-43)
    return F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])


def _cross(hv, ha, gv, ga):
    """hv [nv, BT, r], ha [na, BT, r] -> both updated from the UN-updated other (:380-397)."""
    v_, a_ = hv.transpose(0, 1), ha.transpose(0, 1)                    # [BT, n, r]
    a2v = torch.softmax(v_ @ a_.transpose(1, 2), dim=-1) @ a_
    v2a = torch.softmax(a_ @ v_.transpose(1, 2), dim=-1) @ v_
    return hv + gv * a2v.transpose(0, 1), ha + ga * v2a.transpose(0, 1)


def block(v, a, sd, i, cfg: ClipAVEConfig, fusion, drop=None):
    """ResidualAttentionBlock.forward.  v [nv, BT, D], a [na, BT, D].  drop: optional per-(b n) DropPath masks
    {'v': [B*nv], 'a': [B*na]} already divided by keep (the temporal branch is the only DropPath user)."""
    p = f"transformer.resblocks.{i}."
    T = cfg.num_frames
    out = []
    for x, tag, suffix in ((v, "v", ""), (a, "a", "_Audio")):
        n, bt, D = x.shape
        b = bt // T
        # temporal: 'n (b t) d -> t (b n) d'
        xt = x.reshape(n, b, T, D).permute(2, 1, 0, 3).reshape(T, b * n, D)
        xt = _adapter_out(_adapter_hidden(_mha(_ln(xt, sd, p + "ln_1"), sd, p, cfg.heads), sd, p + "T_Adapter" + suffix), sd,
                          p + "T_Adapter" + suffix)
        if drop is not None:
            xt = xt * drop[tag].reshape(1, b * n, 1).to(xt.dtype)
        xt = xt.reshape(T, b, n, D).permute(2, 1, 0, 3).reshape(n, bt, D)
        out.append(x + xt)
    v, a = out
    gv, ga = sd[p + "gate_v"], sd[p + "gate_a"]
    # spatial
    vs, a_s = _mha(_ln(v, sd, p + "ln_1"), sd, p, cfg.heads), _mha(_ln(a, sd, p + "ln_1"), sd, p, cfg.heads)
    hv, ha = _adapter_hidden(vs, sd, p + "S_Adapter"), _adapter_hidden(a_s, sd, p + "S_Adapter_Audio")
    if fusion:
        hv, ha = _cross(hv, ha, gv, ga)
    v = v + vs + _adapter_out(hv, sd, p + "S_Adapter")
    a = a + a_s + _adapter_out(ha, sd, p + "S_Adapter_Audio")
    # joint
    vn, an = _mlp(_ln(v, sd, p + "ln_2"), sd, p), _mlp(_ln(a, sd, p + "ln_2"), sd, p)
    hv, ha = _adapter_hidden(vn, sd, p + "MLP_Adapter"), _adapter_hidden(an, sd, p + "MLP_Adapter_Audio")
    if fusion:
        hv, ha = _cross(hv, ha, gv, ga)
    v = v + vn + _adapter_out(hv, sd, p + "MLP_Adapter")
    a = a + an + _adapter_out(ha, sd, p + "MLP_Adapter_Audio")
    return v, a


def _embed(x, w, cls, pos, temb, ln_w, ln_b, T):
    """conv patch embed -> [BT, N, D] -> class token, positional + temporal embeddings, ln_pre -> [N+1, BT, D]."""
    x = F.conv2d(x, w, None, stride=w.shape[-1])
    x = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)
    x = torch.cat([cls.to(x.dtype) + torch.zeros(x.shape[0], 1, x.shape[-1], dtype=x.dtype), x], dim=1)
    x = x + pos
    bt, n, D = x.shape
    x = (x.reshape(bt // T, T, n, D) + temb.reshape(1, T, 1, D)).reshape(bt, n, D)
    x = _layer_norm(x, ln_w, ln_b)
    return x.permute(1, 0, 2)


def forward(sd, a, v, cfg: ClipAVEConfig, mode="fusion", drop_masks=None, head_keep=None):
    """-> logits [B*T, label_dim].  drop_masks: {layer: {'v': ..., 'a': ...}} (None = DropPath off);
    head_keep: [B*T, 512] dropout keep mask already divided by keep (None = Dropout in eval)."""
    assert mode in ("fusion", "multimodal")
    B, C, T, H, W = v.shape
    vv = v.permute(0, 2, 1, 3, 4).reshape(B * T, C, H, W)
    aa = a.reshape(B * T, 1, a.shape[2], a.shape[3])
    x_v = _embed(vv, sd["conv1.weight"], sd["class_embedding"], sd["positional_embedding"], sd["temporal_embedding"],
                 sd["ln_pre.weight"], sd["ln_pre.bias"], T)
    x_a = _embed(aa, sd["conv1_audio.weight"], sd["class_embedding"], sd["positional_embedding_audio"],
                 sd["temporal_embedding_audio"], sd["ln_pre.weight"], sd["ln_pre.bias"], T)
    for i in range(cfg.layers):
        x_v, x_a = block(x_v, x_a, sd, i, cfg, mode == "fusion", None if drop_masks is None else drop_masks.get(i))
    fv = _ln(x_v.permute(1, 0, 2), sd, "ln_post")[:, 0]
    fa = _ln(x_a.permute(1, 0, 2), sd, "ln_post")[:, 0]
    h = F.linear(torch.cat((fa, fv), dim=-1), sd["mlp_head.0.weight"], sd["mlp_head.0.bias"])
    if head_keep is not None:
        h = h * head_keep
    return F.linear(h, sd["mlp_head.2.weight"], sd["mlp_head.2.bias"])


def soft_ce(logits, target):
    return -(target * torch.log_softmax(logits, dim=-1)).sum(-1).mean()


def train_step(sd, a, v, labels, cfg: ClipAVEConfig, mode="fusion", **kw):
    """forward + soft-target CrossEntropy + backward under the caller's freeze rule -> (logits, loss, {name: grad})."""
    st = OrderedDict()
    for k, t in sd.items():
        st[k] = t.detach().clone().requires_grad_(True) if (t.is_floating_point() and is_trainable(k)) else t
    out = forward(st, a, v, cfg, mode, **kw)
    loss = soft_ce(out, labels)
    loss.backward()
    grads = {k: t.grad for k, t in st.items() if t.requires_grad and t.grad is not None}
    return out.detach(), loss.detach(), grads
