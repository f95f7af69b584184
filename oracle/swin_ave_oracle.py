"""TEST INFRASTRUCTURE ONLY -- the parity oracle for the Swin AVE adapter train step.

A CPU restatement (plain torch, fp32 or fp64, autograd for the gradients) of what the
reference computes on the hot path named by BASELINE.json:north_star:

    SwinTransformer2D_Adapter_New.forward(a, v, 'fusion')      AVE/model/Swin_AVE.py:1565-1599
      PatchEmbed3D.forward                                     AVE/model/Swin_AVE.py:1104-1124
      SwinTransformerBlock.forward, mode 'fusion_adapt'        AVE/model/Swin_AVE.py:693-813
      WindowAttention.forward (spatial and temporal)           AVE/model/Swin_AVE.py:231-276
      PatchMerging.forward                                     AVE/model/Swin_AVE.py:960-981
      head: norm / avgpool / mlp_head                          AVE/model/Swin_AVE.py:1585-1597
    and mode 'multimodal' (= the same block with the cross-modal term skipped, :490-590).

It is written functionally over a flat ``state`` dict that uses the reference's own
state_dict names, so the same weights drive the reference, this oracle and the CUDA path.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product (`stg-cma_b200/`) never does.

PARITY PINNING: the reference ships no tests, golden vectors or checkpoints (SURVEY.md §4,
§8c), so this restatement is pinned against the reference ITSELF: `oracle/validate_oracle.py`
imports the unmodified reference through `oracle/shim.py` in the build container and checks
logits + every trainable gradient (result recorded in DESIGN.md), and
`oracle/make_goldens.py` stores reference outputs as fixtures under `tests/golden/` which
`tests/test_oracle_golden.py` replays against this file on any machine.
"""
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import List, Optional

import math
import torch
import torch.nn.functional as F


@dataclass
class SwinAVEConfig:
    label_dim: int = 29
    img_size: int = 224
    num_frames: int = 10
    embed_dim: int = 96
    depths: List[int] = field(default_factory=lambda: [2, 2, 6, 2])
    num_heads: List[int] = field(default_factory=lambda: [3, 6, 12, 24])
    window_size: int = 7
    mlp_ratio: float = 4.0
    adapter_mlp_ratio: List[float] = field(default_factory=lambda: [0.125, 0.125, 0.0625, 0.0625])
    patch: int = 4

    @property
    def num_stages(self):
        return len(self.depths)

    def stage_dim(self, s):
        return self.embed_dim * (2 ** s)

    def stage_res(self, s):
        return (self.img_size // self.patch) // (2 ** s)

    def stage_r(self, s):
        # Adapter.__init__: int(D_features * mlp_ratio)          Swin_AVE.py:13
        return int(self.stage_dim(s) * self.adapter_mlp_ratio[s])

    def block_window_shift(self, s, i):
        # SwinTransformerBlock.__init__ (:330-333) + BasicLayer (:1030-1031)
        R = self.stage_res(s)
        ws = self.window_size
        shift = 0 if i % 2 == 0 else ws // 2
        if R <= ws:
            shift = 0
            ws = R
        return ws, shift


PRESETS = {
    # Swin-T: BASELINE.json configs[0]; Swin-B / Swin-L: run_adapt_ave29.py:156-181
    "swin_t": dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                   adapter_mlp_ratio=[0.125, 0.125, 0.0625, 0.0625]),
    "swin_b": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32],
                   adapter_mlp_ratio=[0.125, 0.125, 0.0625, 0.0625]),
    "swin_l": dict(embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48],
                   adapter_mlp_ratio=[0.5, 0.25, 0.125, 0.0625]),
}


# ----------------------------------------------------------------------------------------
# index helpers (integer work: must be bit-exact with the reference's buffers)
# ----------------------------------------------------------------------------------------
def relative_position_index(ws):
    """[ws*ws, ws*ws] int64; idx(i,j) = (yi-yj+ws-1)*(2ws-1) + (xi-xj+ws-1).  Swin_AVE.py:191-200"""
    ys = torch.arange(ws).repeat_interleave(ws)
    xs = torch.arange(ws).repeat(ws)
    dy = ys[:, None] - ys[None, :] + ws - 1
    dx = xs[:, None] - xs[None, :] + ws - 1
    return dy * (2 * ws - 1) + dx


def temporal_relative_coords(T):
    """[T*T] int64; (i - j + T - 1) flattened row-major.  Swin_AVE.py:216-219"""
    t = torch.arange(T)
    return (t[:, None] - t[None, :] + T - 1).reshape(-1)


def shift_region(u, R, ws, shift):
    """Region id of shifted coordinate u along one axis.  Swin_AVE.py:372-377"""
    return (u >= R - ws).long() + (u >= R - shift).long()


def shifted_window_mask(R, ws, shift):
    """[nW, ws*ws, ws*ws] float mask of 0 / -100.0 in shifted coordinates.  Swin_AVE.py:368-387"""
    u = torch.arange(R)
    reg1 = shift_region(u, R, ws, shift)
    reg = reg1[:, None] * 3 + reg1[None, :]                      # [R, R]
    nw = R // ws
    regw = reg.view(nw, ws, nw, ws).permute(0, 2, 1, 3).reshape(nw * nw, ws * ws)
    diff = regw[:, None, :] != regw[:, :, None]
    return diff.to(torch.float32) * -100.0


def window_token_index(R, ws, shift):
    """[nW, ws*ws] int64: which token (row-major y*R+x of the un-shifted map) sits at each
    window slot after roll(-shift) + window_partition.  Swin_AVE.py:728-740, :130-142"""
    nw = R // ws
    pos = torch.arange(R)
    src = (pos + shift) % R                                      # roll by -shift: out[i] = in[(i+shift)%R]
    tok = src[:, None] * R + src[None, :]                        # [R, R] token id at shifted (y, x)
    return tok.view(nw, ws, nw, ws).permute(0, 2, 1, 3).reshape(nw * nw, ws * ws)


# ----------------------------------------------------------------------------------------
# deterministic synthetic weights / inputs
# ----------------------------------------------------------------------------------------
def make_state(cfg: SwinAVEConfig, seed=0, w_std=0.02, gain=1.0, fc1_gain=1.2, fc2_gain=0.5, gate_lo=0.25, gate_hi=0.75,
               dtype=torch.float32, bf16_exact=False):
    """Seeded synthetic weights under the reference's state_dict names (incl. buffers).

    Unlike the reference's init (Swin_AVE.py:1422-1468 zeroes every adapter D_fc2, :365-366
    zero gates) the adapters' second projections and the gates are NON-zero, otherwise
    dW(D_fc1), dgate and the whole cross-modal path are identically zero.

    Every Linear is drawn with std = gain / sqrt(fan_in), so activations stay O(1) at every width
    (Swin-T ... Swin-L) like in a trained network.  A fixed std (e.g. 0.05) makes the adapter
    features grow with the stage width until the UNSCALED cross-modal logits hv.ha^T (Swin_AVE.py:753,
    :801) reach |S| > 100 at stage 3: a softmax that sharp amplifies any input rounding (bf16: 2^-9
    relative -> delta S ~ 0.4 -> tens of percent in P), i.e. an ill-conditioned test problem that no
    reduced-precision implementation -- the reference's own fp16 autocast included -- can meet.  With the
    defaults below the logits are O(1..10) and the cross-modal term still moves the logits by several
    percent (tests/test_oracle_golden.py checks both).

    bf16_exact=True rounds every floating-point tensor to a bf16-representable value (still stored in
    `dtype`): the reference / oracle and the bf16 production path then consume IDENTICAL weights, so a
    parity test measures the kernels and the activation-precision policy, not the (coherent, therefore
    non-averaging) perturbation that quantising given fp32 weights to bf16 would add on top.
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()

    def rn(*shape, std=w_std):
        return (torch.randn(*shape, generator=g, dtype=torch.float64) * std).to(dtype)

    def ln(prefix, n):
        sd[prefix + ".weight"] = (1.0 + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)).to(dtype)
        sd[prefix + ".bias"] = rn(n, std=0.05)

    def linear(prefix, out_f, in_f, bias=True, g_=gain):
        sd[prefix + ".weight"] = rn(out_f, in_f, std=g_ / math.sqrt(in_f))
        if bias:
            sd[prefix + ".bias"] = rn(out_f, std=0.02)

    C0 = cfg.embed_dim
    T = cfg.num_frames
    for name, cin in (("patch_embed", 3), ("patch_embed_audio", 1)):
        sd[name + ".proj.weight"] = rn(C0, cin, 1, cfg.patch, cfg.patch, std=0.1)
        sd[name + ".proj.bias"] = rn(C0, std=0.02)
        ln(name + ".norm", C0)

    for s in range(cfg.num_stages):
        C = cfg.stage_dim(s)
        R = cfg.stage_res(s)
        heads = cfg.num_heads[s]
        r = cfg.stage_r(s)
        hid = int(C * cfg.mlp_ratio)
        for i in range(cfg.depths[s]):
            p = f"layers.{s}.blocks.{i}."
            ws, shift = cfg.block_window_shift(s, i)
            t_attn = (i % 2 == 0)

            def adapter(nm):
                linear(p + nm + ".D_fc1", r, C, g_=fc1_gain)
                linear(p + nm + ".D_fc2", C, r, g_=fc2_gain)

            if t_attn:
                adapter("T_Adapter")
                adapter("T_Adapter_Audio")
            ln(p + "norm1", C)
            sd[p + "attn.relative_position_bias_table"] = rn((2 * ws - 1) * (2 * ws - 1), heads, std=0.2)
            sd[p + "attn.relative_position_index"] = relative_position_index(ws)
            linear(p + "attn.qkv", 3 * C, C)
            linear(p + "attn.proj", C, C)
            if t_attn:
                sd[p + "attn.temporal_position_bias_table"] = rn(2 * T - 1, heads, std=0.2)
                sd[p + "attn.t_relative_coords"] = temporal_relative_coords(T)
                sd[p + "attn.temporal_position_bias_table_audio"] = rn(2 * T - 1, heads, std=0.2)
                sd[p + "attn.t_relative_coords_a"] = temporal_relative_coords(T)
            ln(p + "norm2", C)
            linear(p + "mlp.fc1", hid, C)
            linear(p + "mlp.fc2", C, hid)
            adapter("S_Adapter")
            adapter("S_Adapter2")
            adapter("S_Adapter_Audio")
            adapter("S_Adapter2_Audio")
            sd[p + "gate_v"] = (gate_lo + (gate_hi - gate_lo) * torch.rand(1, generator=g, dtype=torch.float64)).to(dtype)
            sd[p + "gate_a"] = (gate_lo + (gate_hi - gate_lo) * torch.rand(1, generator=g, dtype=torch.float64)).to(dtype)
            if shift > 0:
                sd[p + "attn_mask"] = shifted_window_mask(R, ws, shift).to(dtype)
        if s < cfg.num_stages - 1:
            p = f"layers.{s}.downsample."
            linear(p + "reduction", 2 * C, 4 * C, bias=False)
            ln(p + "norm", 4 * C)

    Cf = cfg.stage_dim(cfg.num_stages - 1)
    ln("norm", Cf)
    linear("mlp_head.0", 512, 2 * Cf)
    linear("mlp_head.2", cfg.label_dim, 512)
    if bf16_exact:
        for k_, t_ in sd.items():
            if t_.is_floating_point() and not is_buffer(k_):
                sd[k_] = t_.to(torch.bfloat16).to(dtype)
    return sd


_BUFFER_KEYS = ("relative_position_index", "t_relative_coords", "attn_mask")
_MLP_LIST = {"mlp_head.0.weight", "mlp_head.0.bias", "mlp_head.2.weight", "mlp_head.2.bias"}
_ADAPT_KEYS = ("adapter", "temporal_embedding", "ln_post", "Adapter", "my_tokens", "gate_", "ln_before",
               "temporal_position_bias_table")


def is_buffer(name):
    return any(k in name for k in _BUFFER_KEYS)


def is_trainable(name):
    """The caller's freeze rule, traintest_adapt_ave29.py:38-61 (freeze_base=True)."""
    if is_buffer(name):
        return False
    if name in _MLP_LIST:
        return True
    return any(k in name for k in _ADAPT_KEYS)


def make_inputs(cfg: SwinAVEConfig, batch, seed=0, dtype=torch.float32):
    """Synthetic clip batch: v ~ N(0,1) [B,3,T,H,W]; a ~ N(0,1) [B,T,H,W];
    soft labels = softmax(N(0,1)) [B*T, label_dim]."""
    g = torch.Generator().manual_seed(1000 + seed)
    S = cfg.img_size
    v = torch.randn(batch, 3, cfg.num_frames, S, S, generator=g, dtype=torch.float32).to(dtype)
    a = torch.randn(batch, cfg.num_frames, S, S, generator=g, dtype=torch.float32).to(dtype)
    labels = torch.softmax(torch.randn(batch * cfg.num_frames, cfg.label_dim, generator=g, dtype=torch.float32), -1).to(dtype)
    return a, v, labels


# ----------------------------------------------------------------------------------------
# forward restatement
# ----------------------------------------------------------------------------------------
def _ln(x, sd, prefix, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def patch_embed(img, sd, prefix, patch):
    """img [B, Cin, T, H, W] -> tokens [(B T), (H/4 W/4), C].  Swin_AVE.py:1104-1124
    tok[b*T+t, y*Wp+x, :] = LN(W[:, c, 0, ky, kx] . img[b, c, t, 4y+ky, 4x+kx] + bias)"""
    B, Cin, T, H, W = img.shape
    Hp, Wp = H // patch, W // patch
    w = sd[prefix + ".proj.weight"].reshape(-1, Cin * patch * patch)          # [C, Cin*16] (c, ky, kx)
    x = img.reshape(B, Cin, T, Hp, patch, Wp, patch).permute(0, 2, 3, 5, 1, 4, 6)  # B T Hp Wp Cin ky kx
    x = x.reshape(B * T, Hp * Wp, Cin * patch * patch)
    x = x @ w.t() + sd[prefix + ".proj.bias"]
    return _ln(x, sd, prefix + ".norm")


def _split_heads(qkv, heads):
    """[..., N, 3C] -> q, k, v each [..., heads, N, hd]; channel order [q|k|v] x heads x hd (:238)"""
    *lead, N, C3 = qkv.shape
    C = C3 // 3
    hd = C // heads
    qkv = qkv.reshape(*lead, N, 3, heads, hd)
    q, k, v = qkv.unbind(-3)
    return q.transpose(-2, -3), k.transpose(-2, -3), v.transpose(-2, -3)


def _merge_heads(o):
    *lead, h, N, hd = o.shape
    return o.transpose(-2, -3).reshape(*lead, N, h * hd)


def temporal_attention(x, sd, p, heads, T, table_name):
    """x [(b t), n, c] -> attention over t per (b, n) with the block's OWN norm1/qkv/proj
    and a trainable [2T-1, heads] bias table.  Swin_AVE.py:705-716, :244-255"""
    BT, L, C = x.shape
    B = BT // T
    xt = x.reshape(B, T, L, C).transpose(1, 2)                                # [B, L, T, C]
    qkv = _lin(_ln(xt, sd, p + "norm1"), sd, p + "attn.qkv")
    q, k, v = _split_heads(qkv, heads)                                        # [B, L, h, T, hd]
    hd = C // heads
    s = (q * hd ** -0.5) @ k.transpose(-1, -2)                                # [B, L, h, T, T]
    tbl = sd[p + "attn." + table_name]                                        # [2T-1, h]
    bias = tbl[temporal_relative_coords(T)].reshape(T, T, heads).permute(2, 0, 1)
    o = torch.softmax(s + bias, dim=-1) @ v
    o = _lin(_merge_heads(o), sd, p + "attn.proj")                            # [B, L, T, C]
    return o                                                                  # still (b n) t c order


def window_attention(xn, sd, p, heads, R, ws, shift):
    """xn = norm1(x) [(b t), L, C] -> proj(softmax(q k^T*scale + rpb + mask) v) returned in
    WINDOW order [(b t), nW, ws*ws, C] together with the token index map.  :718-745, :231-276"""
    BT, L, C = xn.shape
    idx = window_token_index(R, ws, shift)                                    # [nW, N]
    xw = xn[:, idx]                                                           # [BT, nW, N, C]
    qkv = _lin(xw, sd, p + "attn.qkv")
    q, k, v = _split_heads(qkv, heads)                                        # [BT, nW, h, N, hd]
    hd = C // heads
    s = (q * hd ** -0.5) @ k.transpose(-1, -2)
    N = ws * ws
    rpb = sd[p + "attn.relative_position_bias_table"][relative_position_index(ws).reshape(-1)]
    s = s + rpb.reshape(N, N, heads).permute(2, 0, 1)
    if shift > 0:
        s = s + shifted_window_mask(R, ws, shift).to(s.dtype)[None, :, None]
    o = torch.softmax(s, dim=-1) @ v
    return _lin(_merge_heads(o), sd, p + "attn.proj"), idx


def cross_modal(hv, ha, gate_v, gate_a):
    """hv, ha [..., N, r]:  hv' = hv + gate_v * softmax(hv ha^T) ha, ha' likewise; both use
    the un-updated features, logits unscaled.  Swin_AVE.py:753-760 / :801-808"""
    s = hv @ ha.transpose(-1, -2)
    a2v = torch.softmax(s, dim=-1) @ ha
    v2a = torch.softmax(s.transpose(-1, -2), dim=-1) @ hv
    return hv + gate_v * a2v, ha + gate_a * v2a


def block_forward(v, a, sd, cfg, s, i, fusion=True, drop_masks=None):
    """One SwinTransformerBlock in 'fusion_adapt' (fusion=True) or
    'multimodal_adapt_no_fusion' (fusion=False) mode."""
    p = f"layers.{s}.blocks.{i}."
    C = cfg.stage_dim(s)
    R = cfg.stage_res(s)
    heads = cfg.num_heads[s]
    T = cfg.num_frames
    ws, shift = cfg.block_window_shift(s, i)
    BT, L, _ = v.shape
    B = BT // T
    gate_v, gate_a = sd[p + "gate_v"], sd[p + "gate_a"]

    # (T) temporal attention + T_Adapter, even blocks only                  :705-716
    if i % 2 == 0:
        outs = []
        for m, (x, tbl, ad) in enumerate(((v, "temporal_position_bias_table", "T_Adapter"),
                                          (a, "temporal_position_bias_table_audio", "T_Adapter_Audio"))):
            r_t = temporal_attention(x, sd, p, heads, T, tbl)                 # [B, L, T, C]
            r_t = _lin(_gelu(_lin(r_t, sd, p + ad + ".D_fc1")), sd, p + ad + ".D_fc2")
            if drop_masks is not None:                                        # DropPath per (b n) row
                r_t = r_t * drop_masks[(s, i)][m].reshape(B, L, 1, 1).to(r_t.dtype)
            outs.append(x + r_t.transpose(1, 2).reshape(BT, L, C))
        v, a = outs

    # (S) window attention + windowed cross-modal adapter                    :718-787
    av, idx = window_attention(_ln(v, sd, p + "norm1"), sd, p, heads, R, ws, shift)
    aa, _ = window_attention(_ln(a, sd, p + "norm1"), sd, p, heads, R, ws, shift)
    hv = _gelu(_lin(av, sd, p + "S_Adapter2.D_fc1"))
    ha = _gelu(_lin(aa, sd, p + "S_Adapter2_Audio.D_fc1"))
    if fusion:
        hv, ha = cross_modal(hv, ha, gate_v, gate_a)
    av = av + _lin(hv, sd, p + "S_Adapter2.D_fc2")
    aa = aa + _lin(ha, sd, p + "S_Adapter2_Audio.D_fc2")
    flat = idx.reshape(-1)
    inv = torch.empty_like(flat)
    inv[flat] = torch.arange(flat.numel())
    v = v + av.reshape(BT, L, C)[:, inv]                                      # window_reverse + roll back
    a = a + aa.reshape(BT, L, C)[:, inv]

    # (G) frozen MLP + global cross-modal adapter                            :790-811
    yv = _lin(_gelu(_lin(_ln(v, sd, p + "norm2"), sd, p + "mlp.fc1")), sd, p + "mlp.fc2")
    ya = _lin(_gelu(_lin(_ln(a, sd, p + "norm2"), sd, p + "mlp.fc1")), sd, p + "mlp.fc2")
    hv = _gelu(_lin(yv, sd, p + "S_Adapter.D_fc1"))
    ha = _gelu(_lin(ya, sd, p + "S_Adapter_Audio.D_fc1"))
    if fusion:
        hv, ha = cross_modal(hv, ha, gate_v, gate_a)
    v = v + yv + _lin(hv, sd, p + "S_Adapter.D_fc2")
    a = a + ya + _lin(ha, sd, p + "S_Adapter_Audio.D_fc2")
    return v, a


def patch_merge(x, sd, p, R):
    """[(b t), R*R, C] -> [(b t), R*R/4, 2C]; concat order (2y,2x),(2y+1,2x),(2y,2x+1),(2y+1,2x+1)
    then LN(4C) and a bias-free Linear.  Swin_AVE.py:969-979"""
    BT, L, C = x.shape
    x = x.reshape(BT, R // 2, 2, R // 2, 2, C)                                # bt, y, dy, x, dx, c
    x = x.permute(0, 1, 3, 4, 2, 5).reshape(BT, L // 4, 4 * C)                # channel blocks ordered (dx, dy)
    return F.linear(_ln(x, sd, p + "norm"), sd[p + "reduction.weight"])


def forward(sd, a, v, cfg: SwinAVEConfig, mode="fusion", drop_masks=None, head_keep=None):
    """logits [(B T), label_dim].  a [B,T,H,W], v [B,3,T,H,W].  Swin_AVE.py:1565-1599.

    drop_masks: optional {(stage, block): [2, B*L] scale factors} reproducing DropPath on the
    temporal residual; head_keep: optional [(B T), 512] Dropout(0.5) keep-mask already scaled
    by 1/keep.  Both None = deterministic (drop_path_rate=0, head dropout in eval)."""
    if mode not in ("fusion", "multimodal"):
        raise ValueError("oracle covers ftmode 'fusion' and 'multimodal' only")
    xv = patch_embed(v, sd, "patch_embed", cfg.patch)
    xa = patch_embed(a.unsqueeze(1), sd, "patch_embed_audio", cfg.patch)
    for s in range(cfg.num_stages):
        for i in range(cfg.depths[s]):
            xv, xa = block_forward(xv, xa, sd, cfg, s, i, fusion=(mode == "fusion"), drop_masks=drop_masks)
        if s < cfg.num_stages - 1:
            p = f"layers.{s}.downsample."
            xv = patch_merge(xv, sd, p, cfg.stage_res(s))
            xa = patch_merge(xa, sd, p, cfg.stage_res(s))
    fv = _ln(xv, sd, "norm").mean(dim=1)
    fa = _ln(xa, sd, "norm").mean(dim=1)
    h = _lin(torch.cat((fa, fv), dim=-1), sd, "mlp_head.0")                   # audio first (:1596)
    if head_keep is not None:
        h = h * head_keep.to(h.dtype)
    return _lin(h, sd, "mlp_head.2")


def soft_ce(logits, labels):
    """nn.CrossEntropyLoss with probability targets (traintest_adapt_ave29.py:113,163)."""
    return -(labels * torch.log_softmax(logits, dim=-1)).sum(-1).mean()


def train_step(sd, a, v, labels, cfg, mode="fusion", drop_masks=None, head_keep=None):
    """Forward + loss + backward with the caller's freeze rule applied.
    Returns (logits, loss, {name: grad}) for every trainable tensor."""
    work = OrderedDict()
    for k, t in sd.items():
        if is_trainable(k):
            work[k] = t.detach().clone().requires_grad_(True)
        else:
            work[k] = t
    logits = forward(work, a, v, cfg, mode, drop_masks, head_keep)
    loss = soft_ce(logits, labels)
    loss.backward()
    grads = OrderedDict((k, t.grad) for k, t in work.items() if is_trainable(k) and t.grad is not None)
    return logits.detach(), loss.detach(), grads
