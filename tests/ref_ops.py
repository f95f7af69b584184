"""Per-entry-point torch references for the C ABI in include/stgcma.h (TEST CODE).

Every function has the signature of the matching wrapper in stg-cma_b200/stgcma/lib.py and writes
its result into the caller's output tensors.  Two uses:
  * `-m gpu` tests run the CUDA kernel and this reference on the same inputs, op by op;
  * `-m "not gpu"` tests install these functions in place of the ctypes wrappers
    (`install_fake_backend`) so the host-side schedule (stgcma/engine.py: buffer wiring, saved
    tensors, gradient bucket, grouped adapters, drop-path) is checked against the oracle on CPU.
Backward references are torch autograd of the forward references.
"""
import ctypes
import math

import torch
import torch.nn.functional as F

from oracle import swin_ave_oracle as O

EPI_LINEAR, EPI_GELU, EPI_DGELU, EPI_MULAUX, EPI_QGELU, EPI_DQGELU, EPI_GELU_G, EPI_QGELU_G = 0, 1, 2, 3, 4, 5, 6, 7


def _view2d(t, rows, cols, ld):
    return torch.as_strided(t, (rows, cols), (ld, 1))


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def _dgelu(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def _c(t):  # compute dtype for references
    return t.to(torch.float64) if t.dtype != torch.float64 else t


# ---------------------------------------------------------------------------------------------
def patch_embed_fwd(img, w, bias, gamma, beta, out, B, Cin, T, H, W, C):
    x = _c(img).reshape(B, Cin, T, H // 4, 4, W // 4, 4).permute(0, 2, 3, 5, 1, 4, 6).reshape(B * T * (H // 4) * (W // 4), Cin * 16)
    y = x @ _c(w).t() + _c(bias)
    y = F.layer_norm(y, (C,), _c(gamma), _c(beta), 1e-5)
    out.reshape(-1, C).copy_(y)


def _merge_index(frames, R, device):
    """source token (flat over frames*R*R) of each (merged row, quadrant q = dx*2+dy)"""
    Rh = R // 2
    f = torch.arange(frames, device=device).view(-1, 1, 1, 1)
    yy = torch.arange(Rh, device=device).view(1, -1, 1, 1)
    xx = torch.arange(Rh, device=device).view(1, 1, -1, 1)
    q = torch.arange(4, device=device).view(1, 1, 1, -1)
    dy, dx = q % 2, q // 2
    return ((f * R + 2 * yy + dy) * R + 2 * xx + dx).reshape(-1, 4)


def _ln_rows(x, rows, C, merge, R):
    if merge:
        C4 = C // 4
        frames = rows // ((R // 2) ** 2)
        idx = _merge_index(frames, R, x.device)
        return x.reshape(-1, C4)[idx.reshape(-1)].reshape(rows, C), idx
    return x.reshape(-1, C)[:rows], None


def layernorm_fwd(x, gamma, beta, y, mean, rstd, rows, C, merge=0, R=0):
    xr, _ = _ln_rows(_c(x), rows, C, merge, R)
    mu = xr.mean(-1)
    var = ((xr - mu[:, None]) ** 2).mean(-1)
    rs = torch.rsqrt(var + 1e-5)
    y.reshape(-1, C)[:rows].copy_((xr - mu[:, None]) * rs[:, None] * _c(gamma) + _c(beta))
    if mean is not None:
        mean.copy_(mu)
    if rstd is not None:
        rstd.copy_(rs)


def layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx, rows, C, merge=0, R=0, dx_copy=None):
    xr, idx = _ln_rows(_c(x), rows, C, merge, R)
    g = _c(dy).reshape(-1, C)[:rows] * _c(gamma)
    xh = (xr - _c(mean)[:, None]) * _c(rstd)[:, None]
    d = _c(rstd)[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    if merge:
        C4 = C // 4
        tmp = torch.zeros(dx.numel() // C4, C4, dtype=d.dtype, device=d.device)
        tmp[idx.reshape(-1)] = d.reshape(-1, C4)
        dx.reshape(-1, C4).copy_(tmp)
    else:
        if dres is not None:
            d = d + _c(dres).reshape(-1, C)[:rows]
        dx.reshape(-1, C)[:rows].copy_(d)
    if dx_copy is not None:   # the same values, rounded once to the activation dtype (GEMM operand copy)
        dx_copy.reshape(-1).copy_(dx.reshape(-1))


def gemm(A, W, out, M, N, K, *, bias=None, res1=None, res2=None, rowscale=None, aux=None, out_pre=None, lda=None,
         ldw=None, ldo=None, groups=1, epi=EPI_LINEAR):
    lda = K if lda is None else lda
    ldw = K if ldw is None else ldw
    ldo = N if ldo is None else ldo
    Mg = M // groups
    a = _c(_view2d(A, M, K, lda)).reshape(groups, Mg, K)
    w = _c(torch.as_strided(W, (groups, N, K), (N * ldw, ldw, 1)))
    v = torch.einsum("gmk,gnk->gmn", a, w)
    if bias is not None:
        v = v + _c(bias).reshape(groups, 1, N)
    v = v.reshape(M, N)
    if rowscale is not None:
        v = v * _c(rowscale).reshape(M, 1)
    o = _view2d(out, M, N, ldo)
    if epi == EPI_LINEAR:
        if res1 is not None:
            v = v + _c(_view2d(res1, M, N, ldo))
        if res2 is not None:
            v = v + _c(_view2d(res2, M, N, ldo))
        o.copy_(v)
    elif epi == EPI_GELU:
        if out_pre is not None:
            _view2d(out_pre, M, N, ldo).copy_(v)
        o.copy_(_gelu(v))
    elif epi == EPI_DGELU:
        o.copy_(v * _dgelu(_c(_view2d(aux, M, N, ldo))))
    elif epi == EPI_QGELU:        # CLIP's QuickGELU, CLIP_AVE.py:41-43
        if out_pre is not None:
            _view2d(out_pre, M, N, ldo).copy_(v)
        o.copy_(v * torch.sigmoid(1.702 * v))
    elif epi == EPI_DQGELU:
        a = _c(_view2d(aux, M, N, ldo))
        sg = torch.sigmoid(1.702 * a)
        o.copy_(v * sg * (1.0 + 1.702 * a * (1.0 - sg)))
    elif epi == EPI_GELU_G:       # derivative-saving forms: out_pre receives the activation's derivative
        if out_pre is not None:
            _view2d(out_pre, M, N, ldo).copy_(_dgelu(v))
        o.copy_(_gelu(v))
    elif epi == EPI_QGELU_G:
        sg = torch.sigmoid(1.702 * v)
        if out_pre is not None:
            _view2d(out_pre, M, N, ldo).copy_(sg * (1.0 + 1.702 * v * (1.0 - sg)))
        o.copy_(v * sg)
    else:
        o.copy_(v * _c(_view2d(aux, M, N, ldo)))


def mlp_fwd(x, gamma, beta, W1, b1, W2, b2, Wa, ba, y, hpre, h, mean, rstd, M, C, H, rp, groups=2):
    xx = _c(x).reshape(M, C)
    mu = xx.mean(1, keepdim=True)
    var = ((xx - mu) ** 2).mean(1, keepdim=True)
    rs = torch.rsqrt(var + 1e-5)
    xn = ((xx - mu) * rs * _c(gamma) + _c(beta)).to(y.dtype)           # the kernel's A operand is in the activation dtype
    hid = _gelu(_c(xn) @ _c(W1).t() + _c(b1)).to(y.dtype)
    yy = (_c(hid) @ _c(W2).t() + _c(b2)).to(y.dtype)
    y.reshape(M, C).copy_(yy)
    Mg = M // groups
    pre = torch.einsum("gmk,gnk->gmn", _c(yy).reshape(groups, Mg, C), _c(Wa).reshape(groups, rp, C)) + _c(ba).reshape(groups, 1, rp)
    hpre.reshape(M, rp).copy_(pre.reshape(M, rp))
    h.reshape(M, rp).copy_(_gelu(pre).reshape(M, rp))
    mean.copy_(mu.reshape(M))
    rstd.copy_(rs.reshape(M))


def mlp_bwd(dhpre, dxb, x, gamma, beta, mean, rstd, Wat, W1, b1, W2t, W1t, dxn, M, C, H, rp, groups=2):
    xn = ((_c(x).reshape(M, C) - _c(mean).reshape(M, 1)) * _c(rstd).reshape(M, 1) * _c(gamma) + _c(beta)).to(dxn.dtype)
    Mg = M // groups
    dy = torch.einsum("gmk,gnk->gmn", _c(dhpre).reshape(groups, Mg, rp), _c(Wat).reshape(groups, C, rp)).reshape(M, C)
    dy = (dy + _c(dxb).reshape(M, C)).to(dxn.dtype)
    fpre = _c(xn) @ _c(W1).t() + _c(b1)
    df = ((_c(dy) @ _c(W2t).t()) * _dgelu(fpre)).to(dxn.dtype)
    dxn.reshape(M, C).copy_(_c(df) @ _c(W1t).t())


def gemm_tn(A, B, dW, M, Ka, Nb, *, lda, ldb, ldo, rowscale=None, csumA=None, csumB=None):
    a = _c(_view2d(A, M, Ka, lda))
    b = _c(_view2d(B, M, Nb, ldb))
    if rowscale is not None:
        rs = _c(rowscale).reshape(M, 1)
        a = a * rs
        bs = b * rs
    else:
        bs = b
    o = _view2d(dW, Ka, Nb, ldo)
    o.add_((a.t() @ b).to(o.dtype))
    if csumA is not None:
        csumA[:Ka].add_(a.sum(0).to(csumA.dtype))
    if csumB is not None:
        csumB[:Nb].add_(bs.sum(0).to(csumB.dtype))


# ---- attention ------------------------------------------------------------------------------
def _window_attn(qkv, rpb, frames, R, ws, shift, heads):
    C3 = qkv.shape[-1]
    C = C3 // 3
    L = R * R
    idx = O.window_token_index(R, ws, shift).to(qkv.device)           # [nW, N]
    x = qkv.reshape(frames, L, C3)[:, idx]                            # [frames, nW, N, 3C]
    q, k, v = O._split_heads(x, heads)
    N = ws * ws
    s = (q * 32 ** -0.5) @ k.transpose(-1, -2)
    bias = rpb[O.relative_position_index(ws).reshape(-1).to(qkv.device)].reshape(N, N, heads).permute(2, 0, 1)
    s = s + bias
    if shift > 0:
        s = s + O.shifted_window_mask(R, ws, shift).to(s)[None, :, None]
    o = O._merge_heads(torch.softmax(s, -1) @ v)                      # [frames, nW, N, C]
    flat = idx.reshape(-1)
    inv = torch.empty_like(flat)
    inv[flat] = torch.arange(flat.numel(), device=flat.device)
    return o.reshape(frames, L, C)[:, inv]


def window_attn_fwd(qkv, rpb, out, frames, R, ws, shift, heads):
    out.reshape(frames, R * R, -1).copy_(_window_attn(_c(qkv), _c(rpb), frames, R, ws, shift, heads))


@torch.enable_grad()
def window_attn_bwd(qkv, rpb, dout, dqkv, frames, R, ws, shift, heads):
    q = _c(qkv).detach().clone().requires_grad_(True)
    o = _window_attn(q, _c(rpb), frames, R, ws, shift, heads)
    (g,) = torch.autograd.grad(o, q, _c(dout).reshape(o.shape))
    dqkv.copy_(g.reshape(dqkv.shape))


def _temporal_attn(qkv, tbl, G, B, T, L, heads):
    C3 = qkv.shape[-1]
    C = C3 // 3
    x = qkv.reshape(G, B, T, L, C3).permute(0, 1, 3, 2, 4)            # [G, B, L, T, 3C]
    q, k, v = O._split_heads(x, heads)                                # [G, B, L, h, T, 32]
    s = (q * 32 ** -0.5) @ k.transpose(-1, -2)
    rel = O.temporal_relative_coords(T).to(qkv.device)
    bias = tbl[:, rel].reshape(G, T, T, heads).permute(0, 3, 1, 2)    # [G, h, T, T]
    s = s + bias[:, None, None]
    o = O._merge_heads(torch.softmax(s, -1) @ v)                      # [G, B, L, T, C]
    return o.permute(0, 1, 3, 2, 4).reshape(G * B * T * L, C)


def temporal_attn_fwd(qkv, tbl, out, G, B, T, L, heads):
    out.reshape(G * B * T * L, -1).copy_(_temporal_attn(_c(qkv), _c(tbl), G, B, T, L, heads))


@torch.enable_grad()
def temporal_attn_bwd(qkv, tbl, dout, dqkv, dtbl, G, B, T, L, heads):
    q = _c(qkv).detach().clone().requires_grad_(True)
    t = _c(tbl).detach().clone().requires_grad_(True)
    o = _temporal_attn(q, t, G, B, T, L, heads)
    gq, gt = torch.autograd.grad(o, (q, t), _c(dout).reshape(o.shape))
    dqkv.copy_(gq.reshape(dqkv.shape))
    dtbl.reshape(t.shape).add_(gt.to(dtbl.dtype))


def _xmodal(hv, ha, gate):
    s = hv @ ha.transpose(-1, -2)
    a2v = torch.softmax(s, -1) @ ha
    v2a = torch.softmax(s.transpose(-1, -2), -1) @ hv
    return hv + gate[0] * a2v, ha + gate[1] * v2a, a2v, v2a, torch.logsumexp(s, -1), torch.logsumexp(s.transpose(-1, -2), -1)


def _xw(h, gate, frames, R, ws, shift, rp):
    L = R * R
    idx = O.window_token_index(R, ws, shift).to(h.device)
    x = h.reshape(2, frames, L, rp)[:, :, idx]                        # [2, frames, nW, N, rp]
    ov, oa, a2v, v2a, _, _ = _xmodal(x[0], x[1], gate)
    flat = idx.reshape(-1)
    inv = torch.empty_like(flat)
    inv[flat] = torch.arange(flat.numel(), device=flat.device)

    def back(t):
        return t.reshape(frames, L, rp)[:, inv]
    return torch.stack((back(ov), back(oa))), torch.stack((back(a2v), back(v2a)))


def xmodal_window_fwd(h, gate, out, attn_out, frames, R, ws, shift, rp, r=None):
    o, a = _xw(_c(h), _c(gate), frames, R, ws, shift, rp)
    out.copy_(o.reshape(out.shape))
    attn_out.copy_(a.reshape(attn_out.shape))


@torch.enable_grad()
def xmodal_window_bwd(h, hpre, gate, attn_out, dout, dh, dgate, frames, R, ws, shift, rp, r=None):
    hh = _c(h).detach().clone().requires_grad_(True)
    gg = _c(gate).detach().clone().requires_grad_(True)
    o, _ = _xw(hh, gg, frames, R, ws, shift, rp)
    gh, ggate = torch.autograd.grad(o, (hh, gg), _c(dout).reshape(o.shape))
    if hpre is not None:
        gh = gh * _dgelu(_c(hpre)).reshape(gh.shape)
    dh.copy_(gh.reshape(dh.shape))
    dgate.add_(ggate.to(dgate.dtype))


def _xg(h, gate, frames, L, rp, La=None):
    """-> (out, attn_out, lse), each the video block followed by the audio block (flattened rows: La may differ from L)."""
    La = L if La is None else La
    h = h.reshape(-1, rp)
    hv, ha = h[:frames * L].reshape(frames, L, rp), h[frames * L:].reshape(frames, La, rp)
    ov, oa, a2v, v2a, lv, la = _xmodal(hv, ha, gate)
    cat = lambda a, b: torch.cat((a.reshape(-1, rp), b.reshape(-1, rp)))
    return cat(ov, oa), cat(a2v, v2a), torch.cat((lv.reshape(-1), la.reshape(-1)))


def xmodal_global_fwd(h, gate, out, attn_out, lse, frames, L, rp, r=None, La=None):
    o, a, l = _xg(_c(h), _c(gate), frames, L, rp, La)
    out.copy_(o.reshape(out.shape))
    attn_out.copy_(a.reshape(attn_out.shape))
    lse.copy_(l.reshape(lse.shape))


@torch.enable_grad()
def xmodal_global_bwd(h, hpre, gate, attn_out, lse, dout, dh, dgate, work, frames, L, rp, r=None, La=None):
    hh = _c(h).detach().clone().requires_grad_(True)
    gg = _c(gate).detach().clone().requires_grad_(True)
    o, _, _ = _xg(hh, gg, frames, L, rp, La)
    gh, ggate = torch.autograd.grad(o, (hh, gg), _c(dout).reshape(o.shape))
    if hpre is not None:
        gh = gh * _dgelu(_c(hpre)).reshape(gh.shape)
    dh.copy_(gh.reshape(dh.shape))
    dgate.add_(ggate.to(dgate.dtype))


def head_pool_fwd(xn, feat, frames, N, C):
    x = _c(xn).reshape(2, frames, N, C).mean(2)
    feat.copy_(torch.cat((x[1], x[0]), -1))


def head_pool_bwd(dfeat, dxn, frames, N, C):
    d = _c(dfeat).reshape(frames, 2, C) / N                           # [:, 0] audio, [:, 1] video
    dxn.reshape(2, frames, N, C).copy_(torch.stack((d[:, 1], d[:, 0]))[:, :, None, :].expand(2, frames, N, C))


# ---- CLIP-ViT variant ---------------------------------------------------------------------------
def _seq_rows(n_seq, seq_len, inner, device):
    """[n_seq, seq_len] row index of token i of sequence s (include/stgcma.h, stgcma_mha_fwd)."""
    s = torch.arange(n_seq, device=device).view(-1, 1)
    i = torch.arange(seq_len, device=device).view(1, -1)
    return (s // inner) * seq_len * inner + (s % inner) + i * inner


def _mha(qkv, n_seq, seq_len, inner, heads, hd):
    D = heads * hd
    rows = n_seq * seq_len
    idx = _seq_rows(n_seq, seq_len, inner, qkv.device)
    x = qkv.reshape(-1, 3 * D)[:rows][idx]                           # [n_seq, seq_len, 3D]
    q, k, v = (t.reshape(n_seq, seq_len, heads, hd).transpose(1, 2) for t in x.split(D, dim=-1))
    s = (q * hd ** -0.5) @ k.transpose(-1, -2)
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n_seq, seq_len, D)
    lse = torch.logsumexp(s, -1).transpose(1, 2)                     # [n_seq, seq_len, heads]
    inv = torch.empty(rows, dtype=torch.long, device=qkv.device)
    inv[idx.reshape(-1)] = torch.arange(rows, device=qkv.device)
    return o.reshape(rows, D)[inv], lse.reshape(rows, heads)[inv]


def mha_fwd(qkv, out, lse, n_seq, seq_len, inner, heads, hd):
    rows = n_seq * seq_len
    o, l = _mha(_c(qkv), n_seq, seq_len, inner, heads, hd)
    out.reshape(-1, heads * hd)[:rows].copy_(o)
    lse.reshape(-1, heads)[:rows].copy_(l)


@torch.enable_grad()
def mha_bwd(qkv, out, lse, dout, dqkv, work, n_seq, seq_len, inner, heads, hd):
    rows, D = n_seq * seq_len, heads * hd
    q = _c(qkv).reshape(-1, 3 * D)[:rows].detach().clone().requires_grad_(True)
    o, _ = _mha(q, n_seq, seq_len, inner, heads, hd)
    (g,) = torch.autograd.grad(o, q, _c(dout).reshape(-1, D)[:rows])
    dqkv.reshape(-1, 3 * D)[:rows].copy_(g)


def im2col(img, out, B, Cin, T, H, W, P, ldo=None):
    Hp, Wp = (H - P) // P + 1, (W - P) // P + 1
    K = Cin * P * P
    ldo = K if ldo is None else ldo
    x = _c(img).reshape(B, Cin, T, H, W)[:, :, :, :Hp * P, :Wp * P].reshape(B, Cin, T, Hp, P, Wp, P)
    x = x.permute(0, 2, 3, 5, 1, 4, 6).reshape(B * T * Hp * Wp, K)
    _view2d(out, x.shape[0], K, ldo).copy_(x)


def clip_embed_fwd(patches, cls, pos, temb, x0, BT, T, N, D):
    p = _c(patches).reshape(BT, N - 1, D)
    x = torch.cat((_c(cls).reshape(1, 1, D).expand(BT, 1, D), p), 1) + _c(pos).reshape(1, N, D)
    x = x.reshape(BT // T, T, N, D) + _c(temb).reshape(1, T, 1, D)
    x0.reshape(BT, N, D).copy_(x.reshape(BT, N, D))


def clip_temb_bwd(dx0, dtemb, B, T, N, D):
    dtemb.reshape(T, D).add_(_c(dx0).reshape(B, T, N, D).sum((0, 2)).to(dtemb.dtype))


def _cls_rows(x, frames, Nv, Na, D):
    x = x.reshape(-1, D)
    return torch.cat((x[:frames * Nv].reshape(frames, Nv, D)[:, 0], x[frames * Nv:].reshape(frames, Na, D)[:, 0]))   # [2*frames, D]


def cls_ln_fwd(x, gamma, beta, feat, mean, rstd, frames, Nv, Na, D):
    xr = _cls_rows(_c(x), frames, Nv, Na, D)
    mu = xr.mean(-1)
    rs = torch.rsqrt(((xr - mu[:, None]) ** 2).mean(-1) + 1e-5)
    y = (xr - mu[:, None]) * rs[:, None] * _c(gamma) + _c(beta)
    feat.reshape(frames, 2 * D).copy_(torch.cat((y[frames:], y[:frames]), -1))
    mean.copy_(mu)
    rstd.copy_(rs)


def cls_ln_bwd(dfeat, x, mean, rstd, gamma, dx, dgamma, dbeta, frames, Nv, Na, D, dx_copy=None):
    xr = _cls_rows(_c(x), frames, Nv, Na, D)
    df = _c(dfeat).reshape(frames, 2 * D)
    dy = torch.cat((df[:, D:], df[:, :D]))                            # video rows then audio rows
    xh = (xr - _c(mean)[:, None]) * _c(rstd)[:, None]
    g = dy * _c(gamma)
    d = _c(rstd)[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True))
    dxf = dx.reshape(-1, D)
    dxf[:frames * Nv].reshape(frames, Nv, D)[:, 0] = d[:frames].to(dx.dtype)
    dxf[frames * Nv:].reshape(frames, Na, D)[:, 0] = d[frames:].to(dx.dtype)
    if dx_copy is not None:
        cf = dx_copy.reshape(-1, D)
        cf[:frames * Nv].reshape(frames, Nv, D)[:, 0] = d[:frames].to(dx_copy.dtype)
        cf[frames * Nv:].reshape(frames, Na, D)[:, 0] = d[frames:].to(dx_copy.dtype)
    dgamma.add_((dy * xh).sum(0).to(dgamma.dtype))
    dbeta.add_(dy.sum(0).to(dbeta.dtype))


# ---- parameter packing (host memory only: used by the CPU fake backend) ------------------------
_DT = {0: torch.float32, 1: torch.bfloat16, 2: torch.float64}
_CT = {torch.float32: ctypes.c_float, torch.float64: ctypes.c_double}


def pack_params(table_dev, n, dtype_code):
    import struct
    import numpy as np
    st = struct.Struct("<QQiiqqii")
    raw = bytes(table_dev.cpu().numpy().tobytes())
    dst_dtype = _DT[dtype_code]
    for e in range(n):
        src, dst, rows, cols, sld, dld, tr, f32 = st.unpack_from(raw, e * st.size)
        s = np.ctypeslib.as_array(ctypes.cast(src, ctypes.POINTER(ctypes.c_float)), shape=((rows - 1) * sld + cols,))
        sv = torch.from_numpy(np.lib.stride_tricks.as_strided(s, (rows, cols), (sld * 4, 4)).copy())
        dt_ = torch.float32 if f32 else dst_dtype
        if dt_ == torch.bfloat16:  # numpy has no bf16: view the destination through torch
            n_el = ((cols - 1) * dld + rows) if tr else ((rows - 1) * dld + cols)
            d = torch.frombuffer((ctypes.c_int16 * n_el).from_address(dst), dtype=torch.bfloat16)
            dv = torch.as_strided(d, (cols, rows) if tr else (rows, cols), (dld, 1))
            dv.copy_(sv.t() if tr else sv)
            continue
        ct = _CT[dt_]
        esz = ctypes.sizeof(ct)
        if tr:
            d = np.ctypeslib.as_array(ctypes.cast(dst, ctypes.POINTER(ct)), shape=((cols - 1) * dld + rows,))
            dv = np.lib.stride_tricks.as_strided(d, (cols, rows), (dld * esz, esz))
            dv[...] = sv.t().to(dt_).numpy()
        else:
            d = np.ctypeslib.as_array(ctypes.cast(dst, ctypes.POINTER(ct)), shape=((rows - 1) * dld + cols,))
            dv = np.lib.stride_tricks.as_strided(d, (rows, cols), (dld * esz, esz))
            dv[...] = sv.to(dt_).numpy()


def install_fake_backend(monkeypatch, lib_module):
    """Route every kernel wrapper of stgcma.lib to the torch references above (CPU tests only)."""
    import sys
    me = sys.modules[__name__]
    for name in ("patch_embed_fwd", "layernorm_fwd", "layernorm_bwd", "gemm", "gemm_tn", "mlp_fwd", "mlp_bwd", "window_attn_fwd",
                 "window_attn_bwd", "temporal_attn_fwd", "temporal_attn_bwd", "xmodal_window_fwd", "xmodal_window_bwd",
                 "xmodal_global_fwd", "xmodal_global_bwd", "head_pool_fwd", "head_pool_bwd", "pack_params", "mha_fwd", "mha_bwd",
                 "im2col", "clip_embed_fwd", "clip_temb_bwd", "cls_ln_fwd", "cls_ln_bwd"):
        monkeypatch.setattr(lib_module, name, getattr(me, name))
    monkeypatch.setattr(lib_module, "require_device", lambda *t: None)
    monkeypatch.setattr(lib_module, "dt", lambda d: {torch.float32: 0, torch.bfloat16: 1, torch.float64: 2}[
        d.dtype if isinstance(d, torch.Tensor) else d])
