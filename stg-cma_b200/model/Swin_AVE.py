"""Drop-in boundary for the reference's `model.Swin_AVE` (AVE/model/Swin_AVE.py).

`SwinTransformer2D_Adapter_New` keeps the reference's constructor keywords
(run_adapt_ave29.py:156-181), `forward(a, v, mode)` signature (traintest_adapt_ave29.py:158-162),
parameter / buffer names and shapes (so the caller's freeze-by-name rule :38-61, its optimizer
groups, `state_dict()` / `load_state_dict(strict=True)` and checkpoints all keep working), but the
module tree below is only a CONTAINER of fp32 master parameters: no sub-module computes anything.
`forward` hands the whole step to `stgcma.engine.SwinAVEEngine`, i.e. to the sm_100a kernels in
libstgcma.so, through one autograd Function whose backward is the frozen-weight-aware schedule.

Out of scope: ftmode 'videoonly' (different adapter wiring, not
audio-visual) and 'audioonly' (broken in the reference, Swin_AVE.py:1521).
"""
import os
import sys

import torch
import torch.nn as nn

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from stgcma.engine import SwinAVEEngine  # noqa: E402
from stgcma.config import SwinAVEConfig  # noqa: E402


def _trunc_normal_(t, std=0.02):
    return nn.init.trunc_normal_(t, std=std)


class _Container(nn.Module):
    """A holder of parameters; compute happens in the fused engine, never here."""

    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is a parameter container; call the top-level model "
                           "(the B200 path runs the whole step in libstgcma)")


class _Bottleneck(_Container):
    """D_fc1 / D_fc2 pair shared by Adapter, SAdapter2 and T_Adapter (Swin_AVE.py:10-58)."""

    def __init__(self, dim, ratio):
        super().__init__()
        hidden = int(dim * ratio)
        self.D_fc1 = nn.Linear(dim, hidden)
        self.D_fc2 = nn.Linear(hidden, dim)


class Adapter(_Bottleneck):
    pass


class SAdapter2(_Bottleneck):
    pass


class T_Adapter(_Bottleneck):
    pass


class Mlp(_Container):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class WindowAttention(_Container):
    """Parameters + index buffers of Swin_AVE.py:162-229."""

    def __init__(self, dim, num_ttokens, window, num_heads, use_temporal):
        super().__init__()
        side = 2 * window - 1
        self.relative_position_bias_table = nn.Parameter(torch.zeros(side * side, num_heads))
        ys = torch.arange(window).repeat_interleave(window)
        xs = torch.arange(window).repeat(window)
        idx = (ys[:, None] - ys[None, :] + window - 1) * side + (xs[:, None] - xs[None, :] + window - 1)
        self.register_buffer("relative_position_index", idx)
        self.qkv = nn.Linear(dim, 3 * dim, bias=True)
        self.proj = nn.Linear(dim, dim)
        _trunc_normal_(self.relative_position_bias_table)
        if use_temporal:
            t = torch.arange(num_ttokens)
            rel = (t[:, None] - t[None, :] + num_ttokens - 1).reshape(-1)
            self.temporal_position_bias_table = nn.Parameter(torch.zeros(2 * num_ttokens - 1, num_heads))
            _trunc_normal_(self.temporal_position_bias_table)
            self.register_buffer("t_relative_coords", rel.clone())
            self.temporal_position_bias_table_audio = nn.Parameter(torch.zeros(2 * num_ttokens - 1, num_heads))
            _trunc_normal_(self.temporal_position_bias_table_audio)
            self.register_buffer("t_relative_coords_a", rel.clone())


def _shift_mask(R, ws, shift):
    """[nW, ws*ws, ws*ws] 0 / -100 mask in shifted coordinates (Swin_AVE.py:368-387)."""
    u = torch.arange(R)
    reg1 = (u >= R - ws).long() + (u >= R - shift).long()
    reg = reg1[:, None] * 3 + reg1[None, :]
    nw = R // ws
    regw = reg.view(nw, ws, nw, ws).permute(0, 2, 1, 3).reshape(nw * nw, ws * ws)
    return (regw[:, None, :] != regw[:, :, None]).to(torch.float32) * -100.0


class SwinTransformerBlock(_Container):
    def __init__(self, cfg, s, i, drop_path_rate):
        super().__init__()
        dim, R, heads = cfg.stage_dim(s), cfg.stage_res(s), cfg.num_heads[s]
        ws, shift = cfg.block_window_shift(s, i)
        ratio = cfg.adapter_mlp_ratio[s]
        self.t_attn = (i % 2 == 0)
        self.window_size, self.shift_size = ws, shift
        self.drop_path_rate = float(drop_path_rate)
        if self.t_attn:
            self.T_Adapter = T_Adapter(dim, ratio)
            self.T_Adapter_Audio = T_Adapter(dim, ratio)
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, cfg.num_frames, ws, heads, use_temporal=self.t_attn)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * cfg.mlp_ratio))
        self.S_Adapter = Adapter(dim, ratio)
        self.S_Adapter2 = SAdapter2(dim, ratio)
        self.S_Adapter_Audio = Adapter(dim, ratio)
        self.S_Adapter2_Audio = SAdapter2(dim, ratio)
        self.gate_v = nn.Parameter(torch.zeros(1))
        self.gate_a = nn.Parameter(torch.zeros(1))
        # kept for checkpoint compatibility only: the kernels derive the mask from coordinates
        self.register_buffer("attn_mask", _shift_mask(R, ws, shift) if shift > 0 else None)


class PatchMerging(_Container):
    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)


class BasicLayer(_Container):
    def __init__(self, cfg, s, dpr):
        super().__init__()
        self.blocks = nn.ModuleList([SwinTransformerBlock(cfg, s, i, dpr[i]) for i in range(cfg.depths[s])])
        self.downsample = PatchMerging(cfg.stage_dim(s)) if s < cfg.num_stages - 1 else None


class PatchEmbed3D(_Container):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=tuple(patch_size), stride=tuple(patch_size))
        self.norm = nn.LayerNorm(embed_dim)


class _StepFn(torch.autograd.Function):
    """One autograd node for the whole model: forward = engine.forward, backward = engine.backward."""

    @staticmethod
    def forward(ctx, engine, a, v, mode, need_grad, *trainable):
        logits, saved = engine.forward(a, v, mode, need_grad)
        ctx.engine, ctx.saved = engine, saved
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        with torch.no_grad():
            grads = ctx.engine.backward(ctx.saved, dlogits)
        return (None, None, None, None, None, *grads)


class SwinTransformer2D_Adapter_New(nn.Module):
    """Same public surface as the reference class (Swin_AVE.py:1129); see the module docstring.

    Extra keyword (not in the reference): compute_dtype = torch.bfloat16 (production: bf16 storage,
    fp32 accumulate, tensor cores) or torch.float32 (check mode, CUDA cores)."""

    def __init__(self, label_dim, pretrained=None, img_size=224, patch_size=[1, 4, 4], num_frames=10, in_chans=3,
                 embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32], window_size=7, mlp_ratio=4.,
                 frozen_stages=-1, qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2,
                 norm_layer=nn.LayerNorm, ape=False, patch_norm=True, t_relative=True, use_checkpoint=False,
                 ftmode='videoonly', adapter_mlp_ratio=[0.25, 0.25, 0.25, 0.25], compute_dtype=torch.bfloat16, **kwargs):
        super().__init__()
        if ftmode not in ('fusion', 'multimodal'):
            if ftmode in ('videoonly', 'audioonly'):
                raise NotImplementedError(f"ftmode '{ftmode}' is outside the B200 audio-visual hot path "
                                          "(SURVEY.md Appendix C); use 'fusion' or 'multimodal'")
            raise TypeError('ftmode is not expected !!!')
        unsupported = dict(in_chans=(in_chans, 3), qkv_bias=(qkv_bias, True), qk_scale=(qk_scale, None),
                           drop_rate=(drop_rate, 0.), attn_drop_rate=(attn_drop_rate, 0.), ape=(ape, False),
                           patch_norm=(patch_norm, True), t_relative=(t_relative, True), frozen_stages=(frozen_stages, -1))
        for k, (got, want) in unsupported.items():
            if got != want:
                raise NotImplementedError(f"{k}={got!r}: only the reference's default ({want!r}) is on the B200 path")
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("norm_layer must be nn.LayerNorm")
        if list(patch_size) != [1, 4, 4]:
            raise NotImplementedError("patch_size must be [1, 4, 4] (the only value the reference runners use)")
        self.cfg = SwinAVEConfig(label_dim=label_dim, img_size=img_size, num_frames=num_frames, embed_dim=embed_dim,
                                 depths=list(depths), num_heads=list(num_heads), window_size=window_size,
                                 mlp_ratio=mlp_ratio, adapter_mlp_ratio=list(adapter_mlp_ratio))
        cfg = self.cfg
        for s in range(cfg.num_stages):
            if cfg.stage_dim(s) // cfg.num_heads[s] != 32:
                raise NotImplementedError("head_dim must be 32 (every Swin-T/S/B/L configuration)")
        self.num_layers = cfg.num_stages
        self.embed_dim = embed_dim
        self.num_features = cfg.stage_dim(cfg.num_stages - 1)
        self.num_frames = num_frames
        self.patch_size = list(patch_size)
        self.ftmode = ftmode
        self.pretrained = pretrained
        self.compute_dtype = compute_dtype

        self.patch_embed = PatchEmbed3D(patch_size, 3, embed_dim)
        self.patch_embed_audio = PatchEmbed3D(patch_size, 1, embed_dim)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(cfg.depths))]
        self.layers = nn.ModuleList()
        for s in range(cfg.num_stages):
            lo = sum(cfg.depths[:s])
            self.layers.append(BasicLayer(cfg, s, dpr[lo:lo + cfg.depths[s]]))
        self.norm = nn.LayerNorm(self.num_features)
        self.mlp_head = nn.Sequential(nn.Linear(self.num_features * 2, 512), nn.Dropout(0.5), nn.Linear(512, label_dim))
        self.initialize_weights(pretrained)
        self._engine = SwinAVEEngine(self)

    # -- init: trunc-normal(0.02) Linear weights, zero biases, LN (1, 0); every adapter's second
    #    projection and the gates start at zero (Swin_AVE.py:1353-1468), optional Swin checkpoint.
    def initialize_weights(self, pretrained=None):
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                _trunc_normal_(mod.weight)
                if mod.bias is not None:
                    nn.init.zeros_(mod.bias)
            elif isinstance(mod, nn.LayerNorm):
                nn.init.ones_(mod.weight)
                nn.init.zeros_(mod.bias)
        if isinstance(pretrained, str):
            self._load_swin_checkpoint(pretrained)
        elif pretrained is not None:
            raise TypeError('pretrained must be a str or None')
        for name, mod in self.layers.named_modules():
            if isinstance(mod, _Bottleneck):
                nn.init.zeros_(mod.D_fc2.weight)
                nn.init.zeros_(mod.D_fc2.bias)

    def _load_swin_checkpoint(self, path):
        """ImageNet Swin checkpoint -> video + audio streams (Swin_AVE.py:1369-1413): the 2-D patch
        conv is inflated along time, the audio conv is the RGB mean, audio patch-norm = video's."""
        print(f'load model from: {path}')
        sd = torch.load(path, map_location='cpu')['model']
        w = sd['patch_embed.proj.weight']
        pt = self.patch_size[0]
        w3 = w.unsqueeze(2).repeat(1, 1, pt, 1, 1) / pt
        sd['patch_embed.proj.weight'] = w3
        sd['patch_embed_audio.proj.weight'] = w3.mean(dim=1, keepdim=True)
        for k in ('proj.bias', 'norm.weight', 'norm.bias'):
            sd['patch_embed_audio.' + k] = sd['patch_embed.' + k]
        msg = self.load_state_dict(sd, strict=False)
        print('Missing keys: {}'.format(msg.missing_keys))
        print('Unexpected keys: {}'.format(msg.unexpected_keys))
        print(f"=> loaded successfully '{path}'")

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'absolute_pos_embed', 'temporal_embedding', 'temporal_embedding_audio'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table', 'temporal_position_bias_table'}

    def engine(self):
        return self._engine

    def forward(self, a, v, mode):
        if mode not in ('fusion', 'multimodal'):
            raise NotImplementedError(f"mode '{mode}' is outside the B200 audio-visual hot path")
        # the caller wraps this call in fp16 autocast (traintest_adapt_ave29.py:161); the kernels have
        # their own precision policy, so autocast is neutralised here and logits come back fp32.
        with torch.autocast(device_type='cuda', enabled=False):
            trainable = [p for _, p in self._engine.trainable_params()]
            need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in trainable)
            return _StepFn.apply(self._engine, a, v, mode, need_grad, *trainable)
