"""Per-entry-point parity of the CLIP-ViT kernels (stgcma_mha_*, stgcma_im2col, stgcma_clip_embed_fwd,
stgcma_clip_temb_bwd, stgcma_cls_ln_*) against their torch fp64 references in tests/ref_ops.py, through the ctypes
wrappers.  Tolerances as in test_kernels_gpu.py: 1e-4 (fp32 check mode), 1e-2 (bf16) per kernel."""
import pytest
import torch

import ref_ops as R

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16]
TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2}


def L():
    import stgcma.lib as lib
    return lib


def rel(got, ref):
    ref = ref.double().cpu()
    return float((got.double().cpu() - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def rnd(*shape, dtype, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).cuda()


def cpu(t):
    return None if t is None else t.detach().cpu()


MHA_CASES = [
    # n_seq, seq_len, inner, heads, hd
    (6, 197, 1, 2, 96),      # ViT-B/16 video frame: 13 query tiles over 7 warps, ragged last tile
    (9, 49, 1, 8, 96),       # audio frame, all 8 heads, odd sequence count (2 sequences per CTA)
    (2 * 197, 10, 197, 2, 96),   # temporal: T = 10 frames of each (clip, token), rows 197 apart
    (3 * 49, 10, 49, 8, 96),
    (2 * 17, 3, 17, 4, 16),  # the tiny test model: temporal over 3 frames
    (6, 17, 1, 4, 16), (6, 25, 1, 4, 16),
    (3, 257, 1, 2, 64),      # ViT-L/14 video frame (17 query tiles)
    (5, 64, 1, 3, 64),       # ViT-L/14 audio frame
    (3, 70, 1, 2, 32), (2, 100, 1, 1, 128), (1, 1, 1, 1, 16),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", MHA_CASES, ids=lambda c: "s%dx%di%dh%dd%d" % c)
def test_mha(case, dtype):
    n_seq, seq_len, inner, heads, hd = case
    rows, D = n_seq * seq_len, heads * hd
    pad = 5                                   # the kernels run on row-offset views (one modality of the stacked stream)
    qkv_all = rnd(rows + pad, 3 * D, dtype=dtype, seed=1)
    qkv = qkv_all[pad:]
    nan = float("nan")
    out = torch.full((rows, D), nan, dtype=dtype, device="cuda")
    lse = torch.full((rows, heads), nan, device="cuda")
    L().mha_fwd(qkv, out, lse, n_seq, seq_len, inner, heads, hd)
    r_o, r_l = torch.empty(rows, D, dtype=torch.float64), torch.empty(rows, heads, dtype=torch.float64)
    R.mha_fwd(cpu(qkv), r_o, r_l, n_seq, seq_len, inner, heads, hd)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert rel(out, r_o) < TOL[dtype] and rel(lse, r_l) < TOL[dtype]
    dout = rnd(rows, D, dtype=dtype, seed=2)
    dqkv = torch.full((rows, 3 * D), nan, dtype=dtype, device="cuda")
    work = torch.empty(rows, heads, device="cuda")
    L().mha_bwd(qkv, out, lse, dout, dqkv, work, n_seq, seq_len, inner, heads, hd)
    r_d = torch.empty(rows, 3 * D, dtype=torch.float64)
    R.mha_bwd(cpu(qkv), cpu(out), cpu(lse), cpu(dout), r_d, None, n_seq, seq_len, inner, heads, hd)
    assert torch.isfinite(dqkv.float()).all()
    for nm, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        assert rel(dqkv[:, sl], r_d[:, sl]) < TOL[dtype], nm


def test_mha_sharp_rows_bf16():
    """Scores tens of units apart (near one-hot softmax rows), as trained attention produces."""
    n_seq, seq_len, heads, hd = 4, 197, 2, 96
    rows, D = n_seq * seq_len, heads * hd
    qkv = rnd(rows, 3 * D, dtype=torch.bfloat16, seed=5)
    qkv[::5, :D] *= 6
    out = torch.empty(rows, D, dtype=torch.bfloat16, device="cuda")
    lse = torch.empty(rows, heads, device="cuda")
    L().mha_fwd(qkv, out, lse, n_seq, seq_len, 1, heads, hd)
    r_o, r_l = torch.empty(rows, D, dtype=torch.float64), torch.empty(rows, heads, dtype=torch.float64)
    R.mha_fwd(cpu(qkv), r_o, r_l, n_seq, seq_len, 1, heads, hd)
    assert rel(out, r_o) < 1e-2 and rel(lse, r_l) < 1e-2
    dout = rnd(rows, D, dtype=torch.bfloat16, seed=6)
    dqkv = torch.empty(rows, 3 * D, dtype=torch.bfloat16, device="cuda")
    L().mha_bwd(qkv, out, lse, dout, dqkv, torch.empty(rows, heads, device="cuda"), n_seq, seq_len, 1, heads, hd)
    r_d = torch.empty(rows, 3 * D, dtype=torch.float64)
    R.mha_bwd(cpu(qkv), cpu(out), cpu(lse), cpu(dout), r_d, None, n_seq, seq_len, 1, heads, hd)
    assert rel(dqkv, r_d) < 2e-2


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("B,Cin,T,H,W,P,pad", [(2, 3, 2, 64, 64, 16, 0), (1, 1, 3, 102, 128, 16, 0), (2, 1, 2, 48, 128, 16, 8),
                                               (1, 3, 1, 56, 56, 14, 4)])
def test_im2col(B, Cin, T, H, W, P, pad, dtype):
    K = Cin * P * P
    ldo = K + pad
    rows = B * T * ((H - P) // P + 1) * ((W - P) // P + 1)
    img = rnd(B, Cin, T, H, W, dtype=torch.float32, seed=1)
    out = torch.zeros(rows, ldo, dtype=dtype, device="cuda")
    L().im2col(img, out, B, Cin, T, H, W, P, ldo=ldo)
    ref = torch.zeros(rows, ldo, dtype=torch.float64)
    R.im2col(cpu(img), ref, B, Cin, T, H, W, P, ldo=ldo)
    assert rel(out, ref) < (1e-7 if dtype == torch.float32 else 4e-3)
    if pad:
        assert float(out[:, K:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DTYPES)
def test_clip_embed_and_temb_bwd(dtype):
    B, T, N, D = 3, 4, 50, 64
    BT = B * T
    patches = rnd(BT * (N - 1), D, dtype=dtype, seed=1)
    cls, pos, temb = rnd(D, dtype=torch.float32, seed=2), rnd(N, D, dtype=torch.float32, seed=3), rnd(T, D, dtype=torch.float32, seed=4)
    x0 = torch.full((BT * N, D), float("nan"), device="cuda")
    L().clip_embed_fwd(patches, cls, pos, temb, x0, BT, T, N, D)
    ref = torch.empty(BT * N, D, dtype=torch.float64)
    R.clip_embed_fwd(cpu(patches), cpu(cls), cpu(pos), cpu(temb), ref, BT, T, N, D)
    assert rel(x0, ref) < 1e-6
    dx0 = rnd(BT * N, D, dtype=torch.float32, seed=5)
    dtemb = torch.zeros(T, D, device="cuda")
    L().clip_temb_bwd(dx0, dtemb, B, T, N, D)
    r_dt = torch.zeros(T, D, dtype=torch.float64)
    R.clip_temb_bwd(cpu(dx0), r_dt, B, T, N, D)
    assert rel(dtemb, r_dt) < 1e-5


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("frames,Nv,Na,D", [(5, 17, 25, 64), (20, 197, 49, 768)])
def test_cls_ln(frames, Nv, Na, D, dtype):
    rows = frames * (Nv + Na)
    x = rnd(rows, D, dtype=torch.float32, seed=1, scale=2.0) + 0.5
    gamma, beta = 1.0 + 0.1 * rnd(D, dtype=torch.float32, seed=2), 0.1 * rnd(D, dtype=torch.float32, seed=3)
    feat = torch.full((frames, 2 * D), float("nan"), device="cuda")
    mean, rstd = torch.empty(2 * frames, device="cuda"), torch.empty(2 * frames, device="cuda")
    L().cls_ln_fwd(x, gamma, beta, feat, mean, rstd, frames, Nv, Na, D)
    r_f, r_m, r_r = torch.empty(frames, 2 * D, dtype=torch.float64), torch.empty(2 * frames, dtype=torch.float64), torch.empty(2 * frames, dtype=torch.float64)
    R.cls_ln_fwd(cpu(x), cpu(gamma), cpu(beta), r_f, r_m, r_r, frames, Nv, Na, D)
    assert rel(feat, r_f) < 1e-5 and rel(mean, r_m) < 1e-5 and rel(rstd, r_r) < 1e-5
    dfeat = rnd(frames, 2 * D, dtype=torch.float32, seed=4)
    dx, dxc = torch.zeros(rows, D, device="cuda"), torch.zeros(rows, D, dtype=dtype, device="cuda")
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    L().cls_ln_bwd(dfeat, x, mean, rstd, gamma, dx, dg, db, frames, Nv, Na, D, dx_copy=dxc)
    r_dx, r_dg, r_db = torch.zeros(rows, D, dtype=torch.float64), torch.zeros(D, dtype=torch.float64), torch.zeros(D, dtype=torch.float64)
    R.cls_ln_bwd(cpu(dfeat), cpu(x), cpu(mean), cpu(rstd), cpu(gamma), r_dx, r_dg, r_db, frames, Nv, Na, D)
    assert rel(dx, r_dx) < 1e-5 and rel(dg, r_dg) < 1e-5 and rel(db, r_db) < 1e-5
    assert rel(dxc, r_dx) < (1e-5 if dtype == torch.float32 else 4e-3)


def test_bad_arguments_fail_loudly():
    lib = L()
    q = torch.zeros(10, 3 * 24, dtype=torch.bfloat16, device="cuda")
    o, l = torch.zeros(10, 24, dtype=torch.bfloat16, device="cuda"), torch.zeros(10, 1, device="cuda")
    with pytest.raises(lib.StgcmaError):      # head_dim 24 has no tensor-core instantiation
        lib.mha_fwd(q, o, l, 1, 10, 1, 1, 24)
    with pytest.raises(lib.StgcmaError):      # n_seq not a multiple of inner
        lib.mha_fwd(q, o, l, 3, 2, 2, 1, 16)
