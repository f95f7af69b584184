"""Per-entry-point parity: each C-ABI kernel of libstgcma.so (called through the ctypes wrappers of
stgcma/lib.py) against its torch fp64 reference in tests/ref_ops.py on the same seeded inputs.

Tolerances (relative = max|got - ref| / max|ref| per tensor):
  fp32 check mode  : 1e-4 per kernel (the end-to-end bar from BASELINE.json is 1e-3)
  bf16 production  : 1e-2 per kernel (end-to-end bar 2e-2); the reference is fed the SAME
                     bf16-rounded inputs, so this measures the kernel, not the input rounding.
"""
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


# ---------------------------------------------------------------------------------------------
GEMM_CASES = [
    # M, N, K, groups, epi, opts
    (300, 288, 96, 1, R.EPI_LINEAR, dict(bias=True)),                 # K tail (96 = 64 + 32), M tail, BN = 144
    (256, 96, 384, 1, R.EPI_LINEAR, dict(bias=True, res1=True)),
    (1000, 1536, 768, 1, R.EPI_LINEAR, dict(bias=True)),              # several N tiles, 12 k-blocks
    (260, 16, 96, 2, R.EPI_GELU, dict(bias=True, out_pre=True)),      # adapter fc1: grouped, N = 16
    (260, 96, 16, 2, R.EPI_LINEAR, dict(bias=True, res1=True, res2=True, rowscale=True)),  # adapter fc2: K = 16
    (384, 384, 96, 1, R.EPI_GELU, dict(bias=True, out_pre=True)),     # mlp.fc1
    (384, 384, 96, 1, R.EPI_DGELU, dict(aux=True)),
    (130, 48, 192, 2, R.EPI_DGELU, dict(aux=True, rowscale=True)),
    (20, 512, 1536, 1, R.EPI_MULAUX, dict(bias=True, aux=True)),      # head, tiny M
    (20, 32, 512, 1, R.EPI_LINEAR, dict(bias=True)),
    (640, 768, 3072, 1, R.EPI_LINEAR, dict(bias=True, res1=True)),    # mlp.fc2 at stage 3: 48 k-blocks
    (50000, 288, 96, 1, R.EPI_LINEAR, dict(bias=True)),               # many tiles per CTA (persistent loop, phases)
    # QuickGELU epilogues (first brick of the CLIP-ViT row): mlp.c_fc 768 -> 3072 and its dX
    (394, 3072, 768, 1, R.EPI_QGELU, dict(bias=True, out_pre=True)),
    (394, 768, 3072, 1, R.EPI_DQGELU, dict(aux=True)),
    (200, 256, 64, 1, R.EPI_QGELU, dict(bias=True)),
    # derivative-saving forms: out_pre = activation'(v)
    (394, 3072, 768, 1, R.EPI_QGELU_G, dict(bias=True, out_pre=True)),
    (384, 1536, 384, 1, R.EPI_GELU_G, dict(bias=True, out_pre=True)),
    # skinny-K residual GEMMs (gemm_skinny.cu): adapter D_fc2 + residual adds into the fp32 stream, and the dX of D_fc1
    (2 * 147, 96, 16, 2, R.EPI_LINEAR, dict(bias=True, res1=True, res2=True, f32=True)),     # ragged 16-row slabs per group
    (2 * 1000, 384, 32, 2, R.EPI_LINEAR, dict(bias=True, res1=True, res2=True, f32=True)),
    (2 * 333, 768, 48, 2, R.EPI_LINEAR, dict(bias=True, res1=True, f32=True, rowscale=True)),
    (2 * 100, 128, 16, 2, R.EPI_LINEAR, dict(bias=True, res1=True, res2=True)),              # N % 48 != 0: 64-column blocks
    (2 * 130, 192, 96, 2, R.EPI_LINEAR, dict(res1=True)),                                    # Swin-L adapter width, 6 k-steps
    (2 * 40000, 96, 16, 2, R.EPI_LINEAR, dict(bias=True, res1=True, res2=True, rowscale=True, f32=True)),
]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", GEMM_CASES, ids=lambda c: f"M{c[0]}N{c[1]}K{c[2]}g{c[3]}e{c[4]}")
def test_gemm(case, dtype):
    M, N, K, groups, epi, o = case
    A = rnd(M, K, dtype=dtype, seed=1)
    W = rnd(groups, N, K, dtype=dtype, seed=2, scale=K ** -0.5)
    bias = rnd(groups, N, dtype=torch.float32, seed=3) if o.get("bias") else None
    odt = torch.float32 if o.get("f32") else dtype     # fp32 residual stream: out and res1 are fp32 whatever the activation dtype
    res1 = rnd(M, N, dtype=odt, seed=4) if o.get("res1") else None
    res2 = rnd(M, N, dtype=dtype, seed=5) if o.get("res2") else None
    rowscale = (torch.rand(M, generator=torch.Generator().manual_seed(6)) * 2).cuda() if o.get("rowscale") else None
    aux = rnd(M, N, dtype=dtype, seed=7) if o.get("aux") else None
    out = torch.full((M, N), float("nan"), dtype=odt, device="cuda")
    out_pre = torch.full((M, N), float("nan"), dtype=dtype, device="cuda") if o.get("out_pre") else None
    L().gemm(A, W, out, M, N, K, bias=bias, res1=res1, res2=res2, rowscale=rowscale, aux=aux, out_pre=out_pre,
             groups=groups, epi=epi)
    torch.cuda.synchronize()
    r_out = torch.empty(M, N, dtype=torch.float64)
    r_pre = torch.empty(M, N, dtype=torch.float64) if out_pre is not None else None
    R.gemm(cpu(A), cpu(W), r_out, M, N, K, bias=cpu(bias), res1=cpu(res1), res2=cpu(res2), rowscale=cpu(rowscale),
           aux=cpu(aux), out_pre=r_pre, groups=groups, epi=epi)
    assert torch.isfinite(out.float()).all()
    assert rel(out, r_out) < TOL[dtype]
    if out_pre is not None:
        assert rel(out_pre, r_pre) < TOL[dtype]

# ---------------------------------------------------------------------------------------------
# fused frozen MLP (bf16 only): LN2 -> fc1 -> GELU -> fc2 -> adapter fc1 (+GELU), and its dX backward
MLP_CASES = [
    # M, C, H, rp   (groups = 2)
    (2 * 300, 96, 384, 16),       # Swin-T stage 0: K tail (96 = 64 + 32), ragged last tile per group, 2 CTAs / SM
    (2 * 128, 32, 128, 16),       # the tiny test model
    (2 * 1000, 128, 512, 16),     # Swin-B stage 0
    (2 * 200, 192, 768, 32),      # Swin-T stage 1: 128-wide hidden chunks
    (2 * 260, 192, 768, 96),      # Swin-L stage 0 adapter width: two k-blocks of rp in the backward
    (2 * 130, 256, 1024, 32),     # Swin-B stage 1
    (2 * 20000, 96, 384, 16),     # many tiles per CTA: barrier phases over a long persistent loop
]


def _mlp_inputs(M, C, H, rp):
    dt = torch.bfloat16
    x = rnd(M, C, dtype=torch.float32, seed=1, scale=2.0) + 0.5
    gamma = (1.0 + 0.2 * rnd(C, dtype=torch.float32, seed=2))
    beta = 0.1 * rnd(C, dtype=torch.float32, seed=3)
    W1 = rnd(H, C, dtype=dt, seed=4, scale=C ** -0.5)
    b1 = 0.2 * rnd(H, dtype=torch.float32, seed=5)
    W2 = rnd(C, H, dtype=dt, seed=6, scale=H ** -0.5)
    b2 = 0.2 * rnd(C, dtype=torch.float32, seed=7)
    Wa = rnd(2, rp, C, dtype=dt, seed=8, scale=C ** -0.5)
    ba = 0.2 * rnd(2, rp, dtype=torch.float32, seed=9)
    return x, gamma, beta, W1, b1, W2, b2, Wa, ba


@pytest.mark.parametrize("case", MLP_CASES, ids=lambda c: f"M{c[0]}C{c[1]}H{c[2]}r{c[3]}")
def test_mlp_fwd(case):
    M, C, H, rp = case
    dt = torch.bfloat16
    x, gamma, beta, W1, b1, W2, b2, Wa, ba = _mlp_inputs(M, C, H, rp)
    nan = float("nan")
    y = torch.full((M, C), nan, dtype=dt, device="cuda")
    hpre, h = torch.full((M, rp), nan, dtype=dt, device="cuda"), torch.full((M, rp), nan, dtype=dt, device="cuda")
    mean, rstd = torch.full((M,), nan, device="cuda"), torch.full((M,), nan, device="cuda")
    L().mlp_fwd(x, gamma, beta, W1, b1, W2, b2, Wa, ba, y, hpre, h, mean, rstd, M, C, H, rp)
    torch.cuda.synchronize()
    f64 = torch.float64
    r_y, r_hpre, r_h = torch.empty(M, C, dtype=dt), torch.empty(M, rp, dtype=f64), torch.empty(M, rp, dtype=f64)
    r_mean, r_rstd = torch.empty(M, dtype=f64), torch.empty(M, dtype=f64)
    R.mlp_fwd(cpu(x), cpu(gamma), cpu(beta), cpu(W1), cpu(b1), cpu(W2), cpu(b2), cpu(Wa), cpu(ba), r_y, r_hpre, r_h, r_mean,
              r_rstd, M, C, H, rp)
    for t in (y, hpre, h, mean, rstd):
        assert torch.isfinite(t.float()).all()
    assert rel(mean, r_mean) < 1e-5 and rel(rstd, r_rstd) < 1e-5
    assert rel(y, r_y) < 1e-2
    assert rel(hpre, r_hpre) < 1e-2 and rel(h, r_h) < 1e-2


@pytest.mark.parametrize("case", [c for c in MLP_CASES if c[1] <= 256], ids=lambda c: f"M{c[0]}C{c[1]}H{c[2]}r{c[3]}")
def test_mlp_bwd(case):
    M, C, H, rp = case
    dt = torch.bfloat16
    x, gamma, beta, W1, b1, W2, b2, Wa, ba = _mlp_inputs(M, C, H, rp)
    mu = x.mean(1)
    rs = torch.rsqrt(x.var(1, unbiased=False) + 1e-5)
    dhpre = rnd(M, rp, dtype=dt, seed=11)
    dxb = rnd(M, C, dtype=dt, seed=12)
    Wat = Wa.transpose(1, 2).contiguous()
    W2t, W1t = W2.t().contiguous(), W1.t().contiguous()
    dxn = torch.full((M, C), float("nan"), dtype=dt, device="cuda")
    L().mlp_bwd(dhpre, dxb, x, gamma, beta, mu, rs, Wat, W1, b1, W2t, W1t, dxn, M, C, H, rp)
    torch.cuda.synchronize()
    r = torch.empty(M, C, dtype=dt)       # the reference rounds its intermediates like the kernel (bf16 operands)
    R.mlp_bwd(cpu(dhpre), cpu(dxb), cpu(x), cpu(gamma), cpu(beta), cpu(mu), cpu(rs), cpu(Wat), cpu(W1), cpu(b1), cpu(W2t),
              cpu(W1t), r, M, C, H, rp)
    assert torch.isfinite(dxn.float()).all()
    assert rel(dxn, r) < 1e-2


@pytest.mark.parametrize("f32", [False, True])
def test_gemm_skinny_padded_strides(f32):
    """The skinny residual kernel with lda > K and ldo > N (adapter hidden tensors and arena slices are stored padded),
    M not a multiple of 16 and a K of 5 k-steps; the pad columns of `out` must stay untouched."""
    M, N, K, lda, ldo, groups = 2 * 75, 96, 80, 96, 128, 2
    dtype = torch.bfloat16
    odt = torch.float32 if f32 else dtype
    A = rnd(M, lda, dtype=dtype, seed=1)
    W = rnd(groups, N, K, dtype=dtype, seed=2, scale=K ** -0.5)
    bias = rnd(groups, N, dtype=torch.float32, seed=3)
    res1 = rnd(M, ldo, dtype=odt, seed=4)
    res2 = rnd(M, ldo, dtype=dtype, seed=5)
    out = torch.full((M, ldo), 7.0, dtype=odt, device="cuda")
    L().gemm(A, W, out, M, N, K, bias=bias, res1=res1, res2=res2, lda=lda, ldo=ldo, groups=groups)
    torch.cuda.synchronize()
    Mg = M // groups
    ref = torch.cat([cpu(A)[g * Mg:(g + 1) * Mg, :K].double() @ cpu(W)[g].double().t() + cpu(bias)[g].double() for g in range(groups)])
    ref = ref + cpu(res1)[:, :N].double() + cpu(res2)[:, :N].double()
    assert rel(out[:, :N], ref) < (1e-4 if f32 else 1e-2)
    assert float((out[:, N:] - 7.0).abs().max()) == 0.0


def test_gemm_bf16_padded_strides():
    """lda / ldo larger than the logical width (adapter hidden tensors are stored padded)."""
    M, N, K, lda, ldo = 200, 96, 16, 32, 128
    dtype = torch.bfloat16
    A = rnd(M, lda, dtype=dtype, seed=1)
    W = rnd(N, K, dtype=dtype, seed=2)
    out = torch.zeros(M, ldo, dtype=dtype, device="cuda")
    L().gemm(A, W, out, M, N, K, lda=lda, ldo=ldo)
    torch.cuda.synchronize()
    ref = cpu(A)[:, :K].double() @ cpu(W).double().t()
    assert rel(out[:, :N], ref) < 1e-2
    assert float(out[:, N:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("M,Ka,lda,Nb,ldb,rs,colsum_b", [
    (5000, 16, 16, 96, 96, False, True), (333, 96, 96, 12, 16, True, True), (20, 29, 32, 512, 512, False, True),
    (70000, 48, 48, 768, 768, False, True),
    # the engine's calls (no rowscale, no csumB): the bf16 tensor-core kernel
    (50176, 96, 96, 12, 16, False, False), (50176, 12, 16, 96, 96, False, False), (3137, 384, 384, 24, 32, False, False),
    (3137, 24, 32, 384, 384, False, False), (999, 768, 768, 48, 48, False, False), (999, 48, 48, 768, 768, False, False),
    (100, 29, 32, 512, 512, False, False), (70, 192, 192, 24, 32, False, False),
    # DropPath row scale (T adapter): bf16 path multiplies the A fragments by the bf16-rounded scale
    (50177, 96, 96, 12, 16, True, False), (3137, 384, 384, 24, 32, True, False), (999, 48, 48, 768, 768, True, False),
    (130000, 192, 192, 24, 32, False, False), (130000, 96, 96, 96, 96, False, False)])
def test_gemm_tn(M, Ka, lda, Nb, ldb, rs, colsum_b, dtype):
    A = rnd(M, lda, dtype=dtype, seed=1)
    B = rnd(M, ldb, dtype=dtype, seed=2)
    A[:, Ka:] = 0   # pad columns of the engine's tensors are zero
    B[:, Nb:] = 0
    rowscale = (torch.rand(M, generator=torch.Generator().manual_seed(3)) * 2).cuda() if rs else None
    dW = torch.zeros(Ka, Nb, device="cuda")
    ca = torch.zeros(Ka, device="cuda")
    cb = torch.zeros(Nb, device="cuda") if colsum_b else None
    L().gemm_tn(A, B, dW, M, Ka, Nb, lda=lda, ldb=ldb, ldo=Nb, rowscale=rowscale, csumA=ca, csumB=cb)
    torch.cuda.synchronize()
    r_dW, r_ca, r_cb = torch.zeros(Ka, Nb, dtype=torch.float64), torch.zeros(Ka, dtype=torch.float64), torch.zeros(Nb, dtype=torch.float64)
    R.gemm_tn(cpu(A), cpu(B), r_dW, M, Ka, Nb, lda=lda, ldb=ldb, ldo=Nb, rowscale=cpu(rowscale), csumA=r_ca, csumB=r_cb)
    tol = 1e-4  # fp32 accumulation of exactly-represented inputs in both modes
    if rs and dtype == torch.bfloat16 and not colsum_b:
        tol = 1e-2  # tensor-core path: A * rowscale is rounded to bf16 before the product (exact for a 0/1 keep mask)
    assert rel(dW, r_dW) < tol and rel(ca, r_ca) < tol
    if colsum_b:
        assert rel(cb, r_cb) < tol


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("rows,C", [(1000, 96), (77, 384), (64, 1536), (9, 3072)])
def test_layernorm(rows, C, dtype):
    x = rnd(rows, C, dtype=dtype, seed=1, scale=2.0) + 0.5
    gamma, beta = rnd(C, dtype=torch.float32, seed=2) + 1.0, rnd(C, dtype=torch.float32, seed=3)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    L().layernorm_fwd(x, gamma, beta, y, mean, rstd, rows, C)
    r_y, r_m, r_r = torch.empty(rows, C, dtype=torch.float64), torch.empty(rows, dtype=torch.float64), torch.empty(rows, dtype=torch.float64)
    R.layernorm_fwd(cpu(x), cpu(gamma), cpu(beta), r_y, r_m, r_r, rows, C)
    assert rel(y, r_y) < TOL[dtype] and rel(mean, r_m) < 1e-5 and rel(rstd, r_r) < 1e-5
    dy, dres = rnd(rows, C, dtype=dtype, seed=4), rnd(rows, C, dtype=dtype, seed=5)
    dx = torch.empty_like(x)
    L().layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx, rows, C)
    r_dx = torch.empty(rows, C, dtype=torch.float64)
    R.layernorm_bwd(cpu(dy), cpu(x), r_m, r_r, cpu(gamma), cpu(dres), r_dx, rows, C)
    assert rel(dx, r_dx) < TOL[dtype]
    # in-place residual accumulation (dres aliases dx), as the engine uses it
    dx2 = dres.clone()
    L().layernorm_bwd(dy, x, mean, rstd, gamma, dx2, dx2, rows, C)
    assert torch.equal(dx2, dx)


@pytest.mark.parametrize("merge", [0, 1])
def test_layernorm_fp32_stream_bf16_activations(merge):
    """Production layout: x / dres / dx are the fp32 residual stream, y / dy / dx_copy are bf16."""
    frames, Rr, C4 = 3, 14, 96
    C = 4 * C4 if merge else 192
    rows = frames * (Rr // 2) ** 2 if merge else 500
    xshape = (frames * Rr * Rr, C4) if merge else (rows, C)
    x = rnd(*xshape, dtype=torch.float32, seed=1, scale=2.0) + 0.5
    gamma, beta = rnd(C, dtype=torch.float32, seed=2) + 1.0, rnd(C, dtype=torch.float32, seed=3)
    y = torch.empty(rows, C, dtype=torch.bfloat16, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    L().layernorm_fwd(x, gamma, beta, y, mean, rstd, rows, C, merge=merge, R=Rr)
    r_y, r_m, r_r = torch.empty(rows, C, dtype=torch.float64), torch.empty(rows, dtype=torch.float64), torch.empty(rows, dtype=torch.float64)
    R.layernorm_fwd(cpu(x), cpu(gamma), cpu(beta), r_y, r_m, r_r, rows, C, merge=merge, R=Rr)
    assert rel(y, r_y) < 1e-2 and rel(mean, r_m) < 1e-5 and rel(rstd, r_r) < 1e-5
    dy = rnd(rows, C, dtype=torch.bfloat16, seed=4)
    dres = None if merge else rnd(rows, C, dtype=torch.float32, seed=5)
    dx, dxc = torch.full_like(x, float("nan")), torch.full(xshape, float("nan"), dtype=torch.bfloat16, device="cuda")
    L().layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx, rows, C, merge=merge, R=Rr, dx_copy=dxc)
    r_dx = torch.empty(*xshape, dtype=torch.float64)
    R.layernorm_bwd(cpu(dy), cpu(x), r_m, r_r, cpu(gamma), cpu(dres), r_dx, rows, C, merge=merge, R=Rr)
    assert dx.dtype == torch.float32 and rel(dx, r_dx) < 1e-5      # fp32 stream: no bf16 rounding on it
    assert torch.equal(dxc, dx.to(torch.bfloat16))                  # the operand copy is the same value rounded once
    # head variant: fp32 dy as well
    dy32 = dy.float()
    dx2 = torch.empty_like(x)
    L().layernorm_bwd(dy32, x, mean, rstd, gamma, dres, dx2, rows, C, merge=merge, R=Rr, dx_copy=dxc)
    assert torch.equal(dx2, dx)


def test_gemm_bf16_fp32_stream_epilogue():
    """x_new(fp32) = x(fp32) + branch(bf16) + A.W^T + b : `out` / `res1` fp32, everything else bf16."""
    M, N, K = 777, 192, 32
    A, W = rnd(M, K, dtype=torch.bfloat16, seed=1), rnd(2, N, K, dtype=torch.bfloat16, seed=2, scale=K ** -0.5)
    M2 = M - M % 2
    bias = rnd(2, N, dtype=torch.float32, seed=3)
    res1, res2 = rnd(M2, N, dtype=torch.float32, seed=4, scale=30.0), rnd(M2, N, dtype=torch.bfloat16, seed=5)
    out = torch.full((M2, N), float("nan"), dtype=torch.float32, device="cuda")
    L().gemm(A, W, out, M2, N, K, bias=bias, res1=res1, res2=res2, groups=2)
    torch.cuda.synchronize()
    ref = torch.empty(M2, N, dtype=torch.float64)
    R.gemm(cpu(A), cpu(W), ref, M2, N, K, bias=cpu(bias), res1=cpu(res1), res2=cpu(res2), groups=2)
    assert rel(out, ref) < 1e-5   # fp32 out: only fp32 accumulation error, no bf16 rounding of the stream


@pytest.mark.parametrize("dtype", DTYPES)
def test_layernorm_patch_merge(dtype):
    frames, Rr, C4 = 3, 14, 96
    C = 4 * C4
    rows = frames * (Rr // 2) ** 2
    x = rnd(frames * Rr * Rr, C4, dtype=dtype, seed=1)
    gamma, beta = rnd(C, dtype=torch.float32, seed=2) + 1.0, rnd(C, dtype=torch.float32, seed=3)
    y = torch.empty(rows, C, dtype=dtype, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    L().layernorm_fwd(x, gamma, beta, y, mean, rstd, rows, C, merge=1, R=Rr)
    r_y, r_m, r_r = torch.empty(rows, C, dtype=torch.float64), torch.empty(rows, dtype=torch.float64), torch.empty(rows, dtype=torch.float64)
    R.layernorm_fwd(cpu(x), cpu(gamma), cpu(beta), r_y, r_m, r_r, rows, C, merge=1, R=Rr)
    assert rel(y, r_y) < TOL[dtype]
    dy = rnd(rows, C, dtype=dtype, seed=4)
    dx = torch.full_like(x, float("nan"))
    L().layernorm_bwd(dy, x, mean, rstd, gamma, None, dx, rows, C, merge=1, R=Rr)
    r_dx = torch.empty(frames * Rr * Rr, C4, dtype=torch.float64)
    R.layernorm_bwd(cpu(dy), cpu(x), r_m, r_r, cpu(gamma), None, r_dx, rows, C, merge=1, R=Rr)
    assert rel(dx, r_dx) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("Cin,C", [(3, 96), (1, 128), (3, 192)])
def test_patch_embed(Cin, C, dtype):
    B, T, H = 2, 3, 224
    img = rnd(B, Cin, T, H, H, dtype=torch.float32, seed=1)
    w, b = rnd(C, Cin * 16, dtype=torch.float32, seed=2, scale=0.1), rnd(C, dtype=torch.float32, seed=3, scale=0.1)
    gamma, beta = rnd(C, dtype=torch.float32, seed=4) + 1.0, rnd(C, dtype=torch.float32, seed=5)
    out = torch.full((B * T * 56 * 56, C), float("nan"), dtype=dtype, device="cuda")
    L().patch_embed_fwd(img, w, b, gamma, beta, out, B, Cin, T, H, H, C)
    ref = torch.empty(B * T * 56 * 56, C, dtype=torch.float64)
    R.patch_embed_fwd(cpu(img), cpu(w), cpu(b), cpu(gamma), cpu(beta), ref, B, Cin, T, H, H, C)
    assert rel(out, ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("Rr,ws,shift,heads,frames", [(14, 7, 0, 3, 4), (14, 7, 3, 3, 4), (28, 7, 3, 2, 2), (7, 7, 0, 4, 5)])
def test_window_attention(Rr, ws, shift, heads, frames, dtype):
    C = heads * 32
    qkv = rnd(frames * Rr * Rr, 3 * C, dtype=dtype, seed=1)
    rpb = rnd((2 * ws - 1) ** 2, heads, dtype=torch.float32, seed=2, scale=0.5)
    out = torch.full((frames * Rr * Rr, C), float("nan"), dtype=dtype, device="cuda")
    L().window_attn_fwd(qkv, rpb, out, frames, Rr, ws, shift, heads)
    ref = torch.empty(frames * Rr * Rr, C, dtype=torch.float64)
    R.window_attn_fwd(cpu(qkv), cpu(rpb), ref, frames, Rr, ws, shift, heads)
    assert rel(out, ref) < TOL[dtype]
    dout = rnd(frames * Rr * Rr, C, dtype=dtype, seed=3)
    dqkv = torch.full_like(qkv, float("nan"))
    L().window_attn_bwd(qkv, rpb, dout, dqkv, frames, Rr, ws, shift, heads)
    r_d = torch.empty(frames * Rr * Rr, 3 * C, dtype=torch.float64)
    R.window_attn_bwd(cpu(qkv), cpu(rpb), cpu(dout), r_d, frames, Rr, ws, shift, heads)
    assert rel(dqkv, r_d) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("G,B,T,Ln,heads", [(2, 2, 10, 49, 2), (2, 1, 3, 30, 1), (1, 3, 16, 5, 3), (2, 2, 10, 3136, 3),
                                             (2, 1, 10, 196, 12), (2, 1, 10, 50, 24), (1, 1, 2, 7, 1), (2, 3, 10, 797, 6)])
def test_temporal_attention(G, B, T, Ln, heads, dtype):
    C = heads * 32
    rows = G * B * T * Ln
    qkv = rnd(rows, 3 * C, dtype=dtype, seed=1)
    tbl = rnd(G, 2 * T - 1, heads, dtype=torch.float32, seed=2, scale=0.5)
    out = torch.full((rows, C), float("nan"), dtype=dtype, device="cuda")
    L().temporal_attn_fwd(qkv, tbl, out, G, B, T, Ln, heads)
    ref = torch.empty(rows, C, dtype=torch.float64)
    R.temporal_attn_fwd(cpu(qkv), cpu(tbl), ref, G, B, T, Ln, heads)
    assert rel(out, ref) < TOL[dtype]
    dout = rnd(rows, C, dtype=dtype, seed=3)
    dqkv = torch.full_like(qkv, float("nan"))
    dtbl = torch.zeros_like(tbl)
    L().temporal_attn_bwd(qkv, tbl, dout, dqkv, dtbl, G, B, T, Ln, heads)
    r_d, r_t = torch.empty(rows, 3 * C, dtype=torch.float64), torch.zeros(G, 2 * T - 1, heads, dtype=torch.float64)
    R.temporal_attn_bwd(cpu(qkv), cpu(tbl), cpu(dout), r_d, r_t, G, B, T, Ln, heads)
    assert rel(dqkv, r_d) < TOL[dtype]
    assert rel(dtbl, r_t) < max(TOL[dtype], 1e-3)


def _h_inputs(rows2, rp, r, dtype, seed):
    """GELU-like features with the pad columns zero, [2*rows, rp]."""
    h = torch.nn.functional.gelu(rnd(rows2, rp, dtype=torch.float32, seed=seed, scale=1.5)).to(dtype)
    h[:, r:] = 0
    return h


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("frames,Rr,ws,shift,rp,r", [(3, 14, 7, 0, 16, 12), (2, 14, 7, 3, 32, 24), (4, 7, 7, 0, 48, 48),
                                                    (1, 56, 7, 3, 16, 12), (2, 8, 4, 1, 16, 16), (1, 28, 7, 3, 96, 96),
                                                    (1, 14, 7, 0, 64, 64)])
def test_xmodal_window(frames, Rr, ws, shift, rp, r, dtype):
    rows = frames * Rr * Rr
    h = _h_inputs(2 * rows, rp, r, dtype, 1)
    hpre = rnd(2 * rows, rp, dtype=dtype, seed=2)
    gate = torch.tensor([0.6, -0.4], device="cuda")
    out, att = torch.full_like(h, float("nan")), torch.full_like(h, float("nan"))
    L().xmodal_window_fwd(h, gate, out, att, frames, Rr, ws, shift, rp)
    r_o, r_a = torch.empty(2 * rows, rp, dtype=torch.float64), torch.empty(2 * rows, rp, dtype=torch.float64)
    R.xmodal_window_fwd(cpu(h), cpu(gate), r_o, r_a, frames, Rr, ws, shift, rp)
    assert rel(out, r_o) < TOL[dtype] and rel(att, r_a) < TOL[dtype]
    dout = rnd(2 * rows, rp, dtype=dtype, seed=3)
    dh = torch.full_like(h, float("nan"))
    dgate = torch.zeros(2, device="cuda")
    L().xmodal_window_bwd(h, hpre, gate, att, dout, dh, dgate, frames, Rr, ws, shift, rp)
    r_dh, r_dg = torch.empty(2 * rows, rp, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    R.xmodal_window_bwd(cpu(h), cpu(hpre), cpu(gate), cpu(att), cpu(dout), r_dh, r_dg, frames, Rr, ws, shift, rp)
    assert rel(dh, r_dh) < TOL[dtype]
    assert rel(dgate, r_dg) < max(TOL[dtype], 1e-3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("frames,Ln,rp,r,spread", [
    (2, 196, 16, 12, False), (3, 49, 48, 48, False), (1, 100, 96, 96, False), (1, 784, 32, 24, False), (1, 3136, 16, 12, False),
    # rows with 5x larger / 50x smaller features: logsumexp values tens of nats apart, near one-hot softmax rows
    (2, 196, 16, 12, True), (1, 784, 32, 24, True)])
def test_xmodal_global(frames, Ln, rp, r, spread, dtype):
    rows = frames * Ln
    h = _h_inputs(2 * rows, rp, r, dtype, 1)
    if spread:
        h[::7] *= 5
        h[3::50] *= 0.02
    hpre = rnd(2 * rows, rp, dtype=dtype, seed=2)
    gate = torch.tensor([0.7, 0.3], device="cuda")
    out, att = torch.full_like(h, float("nan")), torch.full_like(h, float("nan"))
    lse = torch.empty(2 * rows, device="cuda")
    L().xmodal_global_fwd(h, gate, out, att, lse, frames, Ln, rp)
    r_o, r_a, r_l = (torch.empty(2 * rows, rp, dtype=torch.float64), torch.empty(2 * rows, rp, dtype=torch.float64),
                     torch.empty(2 * rows, dtype=torch.float64))
    R.xmodal_global_fwd(cpu(h), cpu(gate), r_o, r_a, r_l, frames, Ln, rp)
    assert rel(out, r_o) < TOL[dtype] and rel(att, r_a) < TOL[dtype] and rel(lse, r_l) < TOL[dtype]
    dout = rnd(2 * rows, rp, dtype=dtype, seed=3)
    dh = torch.full_like(h, float("nan"))
    dgate = torch.zeros(2, device="cuda")
    work = torch.empty(2 * rows, device="cuda")
    L().xmodal_global_bwd(h, hpre, gate, att, lse, dout, dh, dgate, work, frames, Ln, rp)
    r_dh, r_dg = torch.empty(2 * rows, rp, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    R.xmodal_global_bwd(cpu(h), cpu(hpre), cpu(gate), cpu(att), cpu(lse), cpu(dout), r_dh, r_dg, None, frames, Ln, rp)
    # near one-hot rows amplify the bf16 rounding of P (measured 1.3e-2): the north star's bf16 bar (2e-2) applies
    assert rel(dh, r_dh) < (2e-2 if spread and dtype == torch.bfloat16 else TOL[dtype])
    assert rel(dgate, r_dg) < max(TOL[dtype], 1e-3)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("frames,Lv,La,rp,r", [(3, 197, 49, 48, 48), (2, 49, 197, 16, 12), (1, 130, 7, 32, 24), (2, 64, 65, 16, 16)])
def test_xmodal_global_asymmetric(frames, Lv, La, rp, r, dtype):
    """Different token counts per modality (CLIP ViT-B/16: 197 video / 49 audio tokens per frame, CLIP_AVE.py:380-397):
    rows are the [frames * Lv] video block followed by the [frames * La] audio block."""
    rows = frames * (Lv + La)
    h = _h_inputs(rows, rp, r, dtype, 1)
    hpre = rnd(rows, rp, dtype=dtype, seed=2)
    gate = torch.tensor([0.7, 0.3], device="cuda")
    nan = float("nan")
    out, att = torch.full_like(h, nan), torch.full_like(h, nan)
    lse = torch.full((rows,), nan, device="cuda")
    L().xmodal_global_fwd(h, gate, out, att, lse, frames, Lv, rp, La=La)
    r_o, r_a, r_l = (torch.empty(rows, rp, dtype=torch.float64), torch.empty(rows, rp, dtype=torch.float64),
                     torch.empty(rows, dtype=torch.float64))
    R.xmodal_global_fwd(cpu(h), cpu(gate), r_o, r_a, r_l, frames, Lv, rp, La=La)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert rel(out, r_o) < TOL[dtype] and rel(att, r_a) < TOL[dtype] and rel(lse, r_l) < TOL[dtype]
    dout = rnd(rows, rp, dtype=dtype, seed=3)
    dh = torch.full_like(h, nan)
    dgate = torch.zeros(2, device="cuda")
    work = torch.empty(rows, device="cuda")
    L().xmodal_global_bwd(h, hpre, gate, att, lse, dout, dh, dgate, work, frames, Lv, rp, La=La)
    r_dh, r_dg = torch.empty(rows, rp, dtype=torch.float64), torch.zeros(2, dtype=torch.float64)
    R.xmodal_global_bwd(cpu(h), cpu(hpre), cpu(gate), cpu(att), cpu(lse), cpu(dout), r_dh, r_dg, None, frames, Lv, rp, La=La)
    assert torch.isfinite(dh.float()).all()
    # r = 48 unscaled logits give near one-hot rows: the bf16 rounding of P shows (1.05e-2 measured at 197 x 49): north-star bar
    wide = dtype == torch.bfloat16 and r >= 48
    assert rel(dh, r_dh) < (2e-2 if wide else TOL[dtype])
    assert rel(dgate, r_dg) < (2e-2 if wide else max(TOL[dtype], 1e-3))


@pytest.mark.parametrize("dtype", DTYPES)
def test_head_pool(dtype):
    frames, N, C = 5, 49, 768
    xn = rnd(2 * frames * N, C, dtype=dtype, seed=1)
    feat = torch.empty(frames, 2 * C, dtype=dtype, device="cuda")
    L().head_pool_fwd(xn, feat, frames, N, C)
    ref = torch.empty(frames, 2 * C, dtype=torch.float64)
    R.head_pool_fwd(cpu(xn), ref, frames, N, C)
    assert rel(feat, ref) < TOL[dtype]
    dfeat = rnd(frames, 2 * C, dtype=dtype, seed=2)
    dxn = torch.empty_like(xn)
    L().head_pool_bwd(dfeat, dxn, frames, N, C)
    r_d = torch.empty(2 * frames * N, C, dtype=torch.float64)
    R.head_pool_bwd(cpu(dfeat), r_d, frames, N, C)
    assert rel(dxn, r_d) < TOL[dtype]


@pytest.mark.parametrize("dtype", DTYPES)
def test_pack_params(dtype):
    lib = L()
    src = rnd(12, 96, dtype=torch.float32, seed=1)
    bias = rnd(12, dtype=torch.float32, seed=2)
    w = torch.zeros(16, 96, dtype=dtype, device="cuda")
    wt = torch.zeros(96, 16, dtype=dtype, device="cuda")
    b = torch.zeros(16, device="cuda")
    raw = lib.pack_entry(src, w, 12, 96, 96, 96, False, False) + lib.pack_entry(src, wt, 12, 96, 96, 16, True, False) + \
        lib.pack_entry(bias, b, 1, 12, 12, 16, False, True)
    table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).cuda()
    lib.pack_params(table, 3, lib.dt(dtype))
    torch.cuda.synchronize()
    assert torch.equal(w[:12], src.to(dtype)) and float(w[12:].abs().max()) == 0
    assert torch.equal(wt[:, :12], src.to(dtype).t()) and float(wt[:, 12:].abs().max()) == 0
    assert torch.equal(b[:12], bias) and float(b[12:].abs().max()) == 0


def test_bad_arguments_fail_loudly():
    lib = L()
    x = torch.zeros(8, 100, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(lib.StgcmaError):
        lib.gemm(x, x, x, 8, 100, 100)  # K % 8 != 0 for bf16
    with pytest.raises(lib.StgcmaError):
        lib.xmodal_global_fwd(x, torch.zeros(2, device="cuda"), x, x, torch.zeros(8, device="cuda"), 1, 4, 20)  # rp unsupported


# ---------------------------------------------------------------------------------------------
def test_fused_adam_matches_torch_adam():
    """stgcma_adam_step against torch.optim.Adam with the caller's hyper-parameters (traintest_adapt_ave29.py:68):
    two groups with different learning rates, L2 weight decay, 6 steps, an LR change and a parameter without grad."""
    from stgcma.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(12, 96), (12,), (96, 12), (96,), (1,), (19, 3), (512, 1536), (29, 512), (7,)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    kw = dict(weight_decay=5e-7, betas=(0.95, 0.999))
    mk = lambda ps: [{"params": ps[:6], "lr": 5e-3}, {"params": ps[6:], "lr": 5e-4}]
    ref, ours = torch.optim.Adam(mk(ref_p), **kw), FusedAdam(mk(our_p), **kw)
    for it in range(6):
        if it == 3:
            for o in (ref, ours):
                o.param_groups[0]["lr"] = 1e-3
        for i, (a, b) in enumerate(zip(ref_p, our_p)):
            if i == 8 and it == 3:      # a tensor without a gradient is skipped and keeps its own step count, like torch
                a.grad = b.grad = None
                continue
            gr = torch.randn(*shapes[i], generator=g).cuda() * (10.0 if i == 4 else 1.0)
            a.grad, b.grad = gr.clone(), gr.clone()
        ref.step()
        ours.step()
    torch.cuda.synchronize()
    for a, b in zip(ref_p, our_p):
        assert rel(b.detach(), a.detach()) < 1e-6
    sd = ours.state_dict()
    assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert rel(ours.state[our_p[0]]["exp_avg_sq"], ref.state[ref_p[0]]["exp_avg_sq"]) < 1e-6


def test_fused_adam_under_gradscaler_matches_torch():
    """torch.amp.GradScaler drives FusedAdam through the fused-optimizer protocol (`_step_supports_amp_scaling`): no
    unscale pass, no host read of found_inf -- stgcma_adam_step_amp divides by the device-side scale, skips the update
    when a gradient is inf and keeps the bias corrections on the count of APPLIED steps.  Against torch.optim.Adam under
    its own GradScaler on the same scaled gradients, with an overflow injected in step 2 of 6."""
    from stgcma.optim import FusedAdam
    g = torch.Generator().manual_seed(1)
    shapes = [(12, 96), (96,), (1,), (64, 128), (29, 512)]
    ref_p = [torch.nn.Parameter(torch.randn(*s, generator=g).cuda()) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    kw = dict(weight_decay=5e-7, betas=(0.95, 0.999))
    mk = lambda ps: [{"params": ps[:3], "lr": 5e-3}, {"params": ps[3:], "lr": 5e-4}]
    ref, ours = torch.optim.Adam(mk(ref_p), **kw), FusedAdam(mk(our_p), **kw)
    s_ref, s_our = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for sc in (s_ref, s_our):
        sc.scale(torch.ones(1, device="cuda"))     # creates the device-side scale / growth tracker
    for it in range(6):
        scale = float(s_ref.get_scale())
        assert float(s_our.get_scale()) == scale
        for i, (a, b) in enumerate(zip(ref_p, our_p)):
            gr = torch.randn(*shapes[i], generator=g).cuda() * scale       # what scaler.scale(loss).backward() leaves behind
            if it == 2 and i == 3:
                gr[5, 7] = float("inf")
            a.grad, b.grad = gr.clone(), gr.clone()
        for sc, opt in ((s_ref, ref), (s_our, ours)):
            sc.step(opt)
            sc.update()
    torch.cuda.synchronize()
    assert ours.skipped_steps() == 1
    assert float(s_our.get_scale()) == float(s_ref.get_scale()) == 512.0      # halved once by the overflow
    for a, b in zip(ref_p, our_p):
        assert rel(b.detach(), a.detach()) < 1e-6
    assert rel(ours.state[our_p[3]]["exp_avg_sq"], ref.state[ref_p[3]]["exp_avg_sq"]) < 1e-6
    assert ours.state_dict()["state"][0]["step"] == 5 == int(ref.state_dict()["state"][0]["step"])
    # the gradients are left un-scaled, like torch's fused Adam does
    assert rel(our_p[0].grad, ref_p[0].grad) < 1e-6


def test_fused_adam_load_state_dict_after_step():
    """load_state_dict() on a FusedAdam that has already stepped replaces the moment tensors (and a
    torch.optim.Adam checkpoint stores `step` as a tensor): the device table must follow, and the next step must equal
    torch.optim.Adam's."""
    from stgcma.optim import FusedAdam
    g = torch.Generator().manual_seed(3)
    shapes = [(48, 768), (768,), (1,), (29, 512)]
    base = [torch.randn(s, generator=g) for s in shapes]
    grads = [[torch.randn(s, generator=g) for s in shapes] for _ in range(3)]
    pa = [torch.nn.Parameter(t.clone().cuda()) for t in base]
    pb = [torch.nn.Parameter(t.clone().cuda()) for t in base]
    kw = dict(weight_decay=5e-7, betas=(0.95, 0.999))
    A = FusedAdam([{"params": pa[:2], "lr": 1e-3}, {"params": pa[2:], "lr": 1e-2}], **kw)
    B = torch.optim.Adam([{"params": pb[:2], "lr": 1e-3}, {"params": pb[2:], "lr": 1e-2}], **kw)

    def set_grads(ps, k):
        for p, gr in zip(ps, grads[k]):
            p.grad = gr.clone().cuda()

    set_grads(pa, 0); A.step()                    # A has stepped: its device table exists
    for k in (0, 1):
        set_grads(pb, k); B.step()
    import copy
    A.load_state_dict(copy.deepcopy(B.state_dict()))   # moments + tensor-valued steps (load_state_dict does not copy tensors)
    with torch.no_grad():
        for p, q in zip(pa, pb):
            p.copy_(q)
    set_grads(pa, 2); set_grads(pb, 2)
    A.step(); B.step()
    torch.cuda.synchronize()
    for p, q in zip(pa, pb):
        assert rel(p.detach(), q.detach()) < 1e-6


def test_bucket_allreduce_single_gpu_mechanics():
    """stgcma_bucket_allreduce on ONE GPU: the 'peers' are three local buffers (the kernel only sees a device array of
    pointers), so out = scale * (b0 + b1 + b2) including the n % 4 tail.  The real exchange over peer-mapped memory and
    the NVLS kernels run under torchrun: scripts/dp_exchange_check.py (2 and 8 GPUs, profiles/r02d_dp_exchange_n8.txt)."""
    n = 100003
    bufs = [rnd(n, dtype=torch.float32, seed=40 + i) for i in range(3)]
    ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device="cuda")
    out = torch.full((n,), float("nan"), device="cuda")
    L().bucket_allreduce(ptrs.data_ptr(), 3, out, n, 1.0 / 3.0)
    ref = (bufs[0].double() + bufs[1].double() + bufs[2].double()) / 3.0
    assert rel(out, ref) < 1e-6
    with pytest.raises(L().StgcmaError):
        L().bucket_allreduce(ptrs.data_ptr(), 17, out, n, 1.0)     # more peers than the kernel supports
