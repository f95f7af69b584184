"""Drop-in boundary for the reference's `model.CLIP_AVE` (AVE/model/CLIP_AVE.py; BASELINE.json configs[1]).

`MM_CLIP_AVE` keeps the reference's constructor keywords (CLIP_AVE.py:718-719, call site run_adapt_ave29.py:132-148),
`forward(a, v, mode)` (traintest_adapt_ave29.py:158-162), parameter names and shapes (so the caller's freeze-by-name
rule :38-61, its optimizer groups, `state_dict()` / `load_state_dict(strict=True)` and checkpoints keep working), but the
module tree is only a CONTAINER of fp32 master parameters.  `forward` hands the whole step to
`stgcma.clip_engine.ClipAVEEngine`, i.e. to the sm_100a kernels of libstgcma.so, through one autograd Function whose
backward is the frozen-weight-aware schedule.

In scope: ftmode 'fusion' (block mode 'fusion_adapt', :359-429) and 'multimodal' ('multimodal_adapt_no_fusion',
:256-289).  Out of scope: 'videoonly' / 'audioonly' (single stream, different head).
"""
import os
import sys
from collections import OrderedDict

import torch
import torch.nn as nn

_PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG_ROOT not in sys.path:
    sys.path.insert(0, _PKG_ROOT)

from stgcma.clip_engine import ClipAVEEngine  # noqa: E402
from stgcma.config import ClipAVEConfig  # noqa: E402


class _Container(nn.Module):
    """A holder of parameters; compute happens in the fused engine, never here."""

    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError(f"{type(self).__name__} is a parameter container; call the top-level model "
                           "(the B200 path runs the whole step in libstgcma)")


class Adapter(_Container):
    """D_fc1 / D_fc2 bottleneck, hidden = int(D * mlp_ratio) (CLIP_AVE.py:13-31)."""

    def __init__(self, D_features, mlp_ratio=0.0625, skip_connect=True):
        super().__init__()
        self.skip_connect = skip_connect
        hidden = int(D_features * mlp_ratio)
        self.D_fc1 = nn.Linear(D_features, hidden)
        self.D_fc2 = nn.Linear(hidden, D_features)


class LayerNorm(nn.LayerNorm):
    """Same class name as CLIP_AVE.py:33 (fp32 statistics): here only the affine parameters are used."""


class QuickGELU(_Container):
    pass


class _AttnParams(nn.MultiheadAttention):
    """Holds in_proj_weight / in_proj_bias / out_proj.{weight,bias} under nn.MultiheadAttention's own names and
    initialisation (CLIP_AVE.py:54); never called."""

    def forward(self, *args, **kwargs):  # pragma: no cover
        raise RuntimeError("attention runs inside libstgcma (stgcma_gemm + stgcma_mha_fwd); call the top-level model")


class ResidualAttentionBlock(_Container):
    def __init__(self, d_model, n_head, num_frames, drop_path, ratio):
        super().__init__()
        self.attn = _AttnParams(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = LayerNorm(d_model)
        self.n_head = n_head
        # registration order of the reference (:68-78)
        self.MLP_Adapter = Adapter(d_model, ratio, skip_connect=False)
        self.S_Adapter = Adapter(d_model, ratio)
        self.T_Adapter = Adapter(d_model, ratio, skip_connect=False)
        self.S_Adapter_Audio = Adapter(d_model, ratio)
        self.MLP_Adapter_Audio = Adapter(d_model, ratio, skip_connect=False)
        self.T_Adapter_Audio = Adapter(d_model, ratio, skip_connect=False)
        self.num_frames = num_frames
        self.drop_path_rate = float(drop_path)
        self.gate_v = nn.Parameter(torch.zeros(1))
        self.gate_a = nn.Parameter(torch.zeros(1))


class Transformer(_Container):
    def __init__(self, num_frames, width, layers, heads, drop_path, ratio):
        super().__init__()
        self.width, self.layers = width, layers
        dpr = [x.item() for x in torch.linspace(0, drop_path, layers)]
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, num_frames, dpr[i], ratio) for i in range(layers)])


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


class MM_CLIP_AVE(nn.Module):
    """Same public surface as the reference class (CLIP_AVE.py:716).

    Extra keyword (not in the reference): compute_dtype = torch.bfloat16 (production: bf16 storage, fp32 accumulate,
    tensor cores) or torch.float32 (check mode, CUDA cores)."""

    def __init__(self, label_dim, input_resolution=224, audio_length=1024, num_video_frames=10, patch_size=16, embed_dim=768,
                 layers=12, heads=8, drop_path_rate=0.2, num_tadapter=1, adapter_scale=0.5, pretrained=None,
                 ftmode='videoonly', compute_dtype=torch.bfloat16):
        super().__init__()
        if ftmode not in ('fusion', 'multimodal'):
            if ftmode in ('videoonly', 'audioonly'):
                raise NotImplementedError(f"ftmode '{ftmode}' is outside the B200 audio-visual hot path "
                                          "(SURVEY.md Appendix C); use 'fusion' or 'multimodal'")
            raise TypeError('ftmode is not expected !!!')
        if num_tadapter != 1:
            raise NotImplementedError("num_tadapter must be 1 (the reference's fusion / multimodal blocks use one)")
        if embed_dim % heads or (embed_dim // heads) % 16 or embed_dim // heads > 128:
            raise NotImplementedError("head_dim must be a multiple of 16, at most 128")
        self.ftmode = ftmode
        self.input_resolution = input_resolution
        self.pretrained = pretrained
        self.embed_dim = embed_dim
        self.layers = layers
        self.num_video_frames = num_video_frames
        self.compute_dtype = compute_dtype
        self.cfg = ClipAVEConfig(label_dim=label_dim, input_resolution=input_resolution, audio_length=audio_length,
                                 num_frames=num_video_frames, patch_size=patch_size, embed_dim=embed_dim, layers=layers,
                                 heads=heads)
        cfg = self.cfg
        self.ori_num_patches = cfg.grid_v ** 2
        self.oringal_hw = cfg.grid_v
        # the reference probes the audio patch grid with a [1, 1, 128, audio_length / 10] test convolution (:793-800):
        # f_dim over the 128 mel bins, t_dim over time
        self.f_dim, self.t_dim = self.get_shape_a(fstride=patch_size, tstride=patch_size, input_fdim=128,
                                                  input_tdim=int(audio_length * (1 / 10)))
        self.num_patches_a = self.f_dim * self.t_dim
        if self.num_patches_a != cfg.na - 1:
            raise NotImplementedError(f"audio patch grid {cfg.grid_a} of the {patch_size} x {patch_size} convolution does not match "
                                      f"the reference's probe grid ({self.f_dim}, {self.t_dim}): the reference itself fails here")

        scale = embed_dim ** -0.5
        self.conv1 = nn.Conv2d(3, embed_dim, kernel_size=patch_size, stride=patch_size, bias=False)
        self.class_embedding = nn.Parameter(scale * torch.randn(embed_dim))
        self.positional_embedding = nn.Parameter(scale * torch.randn(cfg.nv, embed_dim))
        self.ln_pre = LayerNorm(embed_dim)
        self.conv1_audio = nn.Conv2d(1, embed_dim, kernel_size=patch_size, stride=patch_size, bias=False)
        self.positional_embedding_audio = nn.Parameter(scale * torch.randn(cfg.na, embed_dim))
        self.temporal_embedding = nn.Parameter(torch.zeros(1, num_video_frames, embed_dim))
        self.temporal_embedding_audio = nn.Parameter(torch.zeros(1, num_video_frames, embed_dim))
        self.transformer = Transformer(num_video_frames, embed_dim, layers, heads, drop_path_rate, cfg.adapter_ratio)
        self.ln_post = LayerNorm(embed_dim)
        self.mlp_head = nn.Sequential(nn.Linear(embed_dim * 2, 512), nn.Dropout(0.5), nn.Linear(512, label_dim))
        self.initialize_weights(pretrained)
        print('Visual Positional Embedding Shape:', self.positional_embedding.shape)
        print('Temporal Positional Embedding Shape:', self.temporal_embedding.shape)
        print('Audio Positional Embedding Shape:', self.positional_embedding_audio.shape)
        self._engine = ClipAVEEngine(self)

    def get_shape_a(self, fstride=16, tstride=16, input_fdim=128, input_tdim=1024):
        """(f_dim, t_dim) of the reference's probe convolution (:793-800): 16 x 16 kernel, the given strides."""
        return (input_fdim - 16) // fstride + 1, (input_tdim - 16) // tstride + 1

    # -- init (CLIP_AVE.py:802-966): trunc-normal(0.02) Linear weights, zero biases, LN (1, 0), every adapter's second
    #    projection zero, optional OpenAI CLIP visual-tower checkpoint.
    def initialize_weights(self, pretrained=None):
        for mod in self.modules():
            if isinstance(mod, nn.Linear):
                nn.init.trunc_normal_(mod.weight, std=.02)
                if mod.bias is not None:
                    nn.init.constant_(mod.bias, 0)
            elif isinstance(mod, nn.LayerNorm):
                nn.init.constant_(mod.bias, 0)
                nn.init.constant_(mod.weight, 1.0)
        if isinstance(pretrained, str):
            self._load_clip_checkpoint(pretrained)
        elif pretrained is not None:
            raise TypeError('pretrained must be a str or None')
        for mod in self.transformer.modules():
            if isinstance(mod, Adapter):
                nn.init.constant_(mod.D_fc2.weight, 0)
                nn.init.constant_(mod.D_fc2.bias, 0)

    def _load_clip_checkpoint(self, path):
        """OpenAI CLIP visual tower -> video + audio streams (CLIP_AVE.py:813-858): `proj` dropped, the audio patch
        conv is the RGB sum, the audio positional embedding is a centre crop (or bilinear resize) of the 2-D grid of the
        video one.  `path` is the reference's download_root (needs the `clip` package), or a file holding the visual
        tower's state_dict (optionally under a 'visual.' prefix) saved with torch.save."""
        print(f'load model from: {path}')
        if os.path.isfile(path):
            sd = torch.load(path, map_location='cpu')
            sd = sd.get('state_dict', sd) if isinstance(sd, dict) else sd.state_dict()
            if any(k.startswith('visual.') for k in sd):
                sd = {k[len('visual.'):]: v for k, v in sd.items() if k.startswith('visual.')}
            pretrain_dict = OrderedDict((k, v.float()) for k, v in sd.items())
        else:
            import clip
            clip_model, _ = clip.load("ViT-B/16" if self.layers == 12 else "ViT-L/14", device="cpu", download_root=path)
            pretrain_dict = clip_model.visual.state_dict()
            del clip_model
        pretrain_dict.pop('proj', None)
        pretrain_dict['conv1_audio.weight'] = torch.sum(pretrain_dict['conv1.weight'], dim=1).unsqueeze(1)
        hw, D = self.oringal_hw, self.embed_dim
        ori = pretrain_dict['positional_embedding'].unsqueeze(0)
        a_pos = ori[:, 1:, :].detach().reshape(1, self.ori_num_patches, D).transpose(1, 2).reshape(1, D, hw, hw)
        if self.t_dim <= hw:
            lo = int(hw / 2) - int(self.t_dim / 2)
            a_pos = a_pos[:, :, :, lo:lo + self.t_dim]
        else:
            a_pos = torch.nn.functional.interpolate(a_pos, size=(hw, self.t_dim), mode='bilinear')
        if self.f_dim <= hw:
            lo = int(hw / 2) - int(self.f_dim / 2)
            a_pos = a_pos[:, :, lo:lo + self.f_dim, :]
        else:
            a_pos = torch.nn.functional.interpolate(a_pos, size=(self.f_dim, self.t_dim), mode='bilinear')
        a_pos = a_pos.reshape(1, D, self.num_patches_a).transpose(1, 2)
        pretrain_dict['positional_embedding_audio'] = torch.cat([ori[:, :1, :].detach(), a_pos], dim=1).squeeze(dim=0)
        msg = self.load_state_dict(pretrain_dict, strict=False)
        print('Missing keys: {}'.format(msg.missing_keys))
        print('Unexpected keys: {}'.format(msg.unexpected_keys))
        print(f"=> loaded successfully '{path}'")

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'absolute_pos_embed', 'temporal_embedding', 'positional_embedding_audio', 'temporal_embedding_audio'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table', 'temporal_position_bias_table'}

    def engine(self):
        return self._engine

    def forward(self, a, v, mode):
        if mode not in ('fusion', 'multimodal'):
            raise NotImplementedError(f"mode '{mode}' is outside the B200 audio-visual hot path")
        # the caller wraps this call in fp16 autocast (traintest_adapt_ave29.py:161); the kernels have their own
        # precision policy, so autocast is neutralised here and logits come back fp32.
        with torch.autocast(device_type='cuda', enabled=False):
            trainable = [p for _, p in self._engine.trainable_params()]
            need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in trainable)
            return _StepFn.apply(self._engine, a, v, mode, need_grad, *trainable)
