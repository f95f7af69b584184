"""Static shape description of the Swin AVE adapter model (reference ctor arguments,
AVE/model/Swin_AVE.py:1155-1160, AVE/run_adapt_ave29.py:156-181)."""
from dataclasses import dataclass, field
from typing import List


@dataclass
class SwinAVEConfig:
    label_dim: int = 29
    img_size: int = 224
    num_frames: int = 10
    embed_dim: int = 128
    depths: List[int] = field(default_factory=lambda: [2, 2, 18, 2])
    num_heads: List[int] = field(default_factory=lambda: [4, 8, 16, 32])
    window_size: int = 7
    mlp_ratio: float = 4.0
    adapter_mlp_ratio: List[float] = field(default_factory=lambda: [0.25, 0.25, 0.25, 0.25])
    patch: int = 4

    @property
    def num_stages(self):
        return len(self.depths)

    def stage_dim(self, s):
        return int(self.embed_dim * 2 ** s)

    def stage_res(self, s):
        return (self.img_size // self.patch) // (2 ** s)

    def stage_r(self, s):
        """adapter hidden width, int(D_features * mlp_ratio) (Swin_AVE.py:13)"""
        return int(self.stage_dim(s) * self.adapter_mlp_ratio[s])

    def block_window_shift(self, s, i):
        """(window, shift) of block i of stage s (Swin_AVE.py:330-333, :1030-1031)"""
        R, ws = self.stage_res(s), self.window_size
        shift = 0 if i % 2 == 0 else ws // 2
        if R <= ws:
            ws, shift = R, 0
        return ws, shift


# named model sizes: Swin-T = BASELINE.json configs[0]; Swin-B / Swin-L = run_adapt_ave29.py:156-181
PRESETS = {
    "swin_t": dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24],
                   adapter_mlp_ratio=[0.125, 0.125, 0.0625, 0.0625]),
    "swin_b": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32],
                   adapter_mlp_ratio=[0.125, 0.125, 0.0625, 0.0625]),
    "swin_l": dict(embed_dim=192, depths=[2, 2, 18, 2], num_heads=[6, 12, 24, 48],
                   adapter_mlp_ratio=[0.5, 0.25, 0.125, 0.0625]),
}


@dataclass
class ClipAVEConfig:
    """Static shape description of `MM_CLIP_AVE` (reference ctor arguments, AVE/model/CLIP_AVE.py:718-719,
    AVE/run_adapt_ave29.py:132-139).  ViT-B/16 as the reference runs it: 12 layers, D = 768, EIGHT heads of 96."""
    label_dim: int = 29
    input_resolution: int = 224
    audio_length: int = 1024        # fbank target_length; the model sees [B, T, audio_length // 10, 128]
    num_frames: int = 10
    patch_size: int = 16
    embed_dim: int = 768
    layers: int = 12
    heads: int = 8
    adapter_ratio: float = 0.0625   # Adapter(mlp_ratio=0.0625), CLIP_AVE.py:14

    @property
    def grid_v(self):
        return self.input_resolution // self.patch_size

    @property
    def nv(self):                   # video tokens per frame incl. the class token
        return self.grid_v ** 2 + 1

    @property
    def audio_shape(self):          # (H, W) of one audio image: [audio_length / 10 time steps, 128 mel bins]
        return int(self.audio_length * (1 / 10)), 128

    @property
    def grid_a(self):               # conv output (rows, cols) on an [H, W] audio image, kernel = stride = patch
        h, w = self.audio_shape
        return (h - self.patch_size) // self.patch_size + 1, (w - self.patch_size) // self.patch_size + 1

    @property
    def na(self):
        gh, gw = self.grid_a
        return gh * gw + 1

    @property
    def r(self):
        return int(self.embed_dim * self.adapter_ratio)

    @property
    def head_dim(self):
        return self.embed_dim // self.heads


CLIP_PRESETS = {
    # BASELINE.json configs[1]: ViT-B/16 exactly as run_adapt_ave29.py:132-139 builds it
    "clip_b16": dict(input_resolution=224, patch_size=16, embed_dim=768, layers=12, heads=8),
    # ViT-L/14 (run_adapt_ave29.py:141-148)
    "clip_l14": dict(input_resolution=224, patch_size=14, embed_dim=1024, layers=24, heads=16),
}
