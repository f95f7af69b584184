"""ctypes binding of libstgcma.so (C ABI: include/stgcma.h).

There is NO fallback: if the shared library is missing, fails to load, or the device is not
sm_100, every entry point raises.  PyTorch is used only for device memory and streams; the
arguments that cross the boundary are raw device pointers, sizes and the current CUDA stream.
"""
import ctypes
import os
import struct

import torch

F32, BF16 = 0, 1
EPI_LINEAR, EPI_GELU, EPI_DGELU, EPI_MULAUX, EPI_QGELU, EPI_DQGELU, EPI_GELU_G, EPI_QGELU_G = 0, 1, 2, 3, 4, 5, 6, 7

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstgcma.so")

_vp, _i, _i64, _fp = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p

# name -> argtypes (restype is always int unless noted)
_SIGNATURES = {
    "stgcma_check_device": [_i],
    "stgcma_patch_embed_fwd": [_fp, _fp, _fp, _fp, _fp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "stgcma_layernorm_fwd": [_vp, _fp, _fp, _vp, _fp, _fp, _i64, _i, _i, _i, _i, _i, _vp],
    "stgcma_layernorm_bwd": [_vp, _vp, _fp, _fp, _fp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp],
    "stgcma_gemm": [_vp, _vp, _fp, _vp, _vp, _fp, _vp, _vp, _vp, _i64, _i, _i, _i64, _i64, _i64, _i, _i, _i, _i, _vp],
    "stgcma_mlp_fwd": [_fp, _fp, _fp, _vp, _fp, _vp, _fp, _vp, _fp, _vp, _vp, _vp, _fp, _fp, _i64, _i, _i, _i, _i, _i, _vp],
    "stgcma_mlp_bwd": [_vp, _vp, _fp, _fp, _fp, _fp, _fp, _vp, _vp, _fp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "stgcma_gemm_tn": [_vp, _vp, _fp, _fp, _fp, _fp, _i64, _i, _i, _i64, _i64, _i64, _i, _vp],
    "stgcma_window_attn_fwd": [_vp, _fp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "stgcma_window_attn_bwd": [_vp, _fp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "stgcma_temporal_attn_fwd": [_vp, _fp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    "stgcma_temporal_attn_bwd": [_vp, _fp, _vp, _vp, _fp, _i, _i, _i, _i, _i, _i, _vp],
    "stgcma_xmodal_window_fwd": [_vp, _fp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "stgcma_xmodal_window_bwd": [_vp, _vp, _fp, _vp, _vp, _vp, _fp, _i64, _i, _i, _i, _i, _i, _vp],
    "stgcma_xmodal_global_fwd": [_vp, _fp, _vp, _vp, _fp, _i64, _i, _i, _i, _i, _vp],
    "stgcma_xmodal_global_bwd": [_vp, _vp, _fp, _vp, _fp, _vp, _vp, _fp, _fp, _i64, _i, _i, _i, _i, _vp],
    "stgcma_head_pool_fwd": [_vp, _vp, _i64, _i, _i, _i, _vp],
    "stgcma_head_pool_bwd": [_vp, _vp, _i64, _i, _i, _i, _vp],
    "stgcma_pack_params": [_vp, _i, _i, _vp],
    "stgcma_adam_step": [_vp, _i, _vp, ctypes.c_float, _vp],
    "stgcma_adam_step_amp": [_vp, _i, _vp, _vp, _fp, _fp, _vp, _vp],
    "stgcma_bucket_allreduce": [_vp, _i, _fp, _fp, _i64, ctypes.c_float, _vp],
    "stgcma_bucket_allreduce_nvls": [_fp, _fp, _i64, _i, _i, ctypes.c_float, _vp],
    "stgcma_mha_fwd": [_vp, _vp, _fp, _i64, _i, _i, _i, _i, _i, _vp],
    "stgcma_mha_bwd": [_vp, _vp, _fp, _vp, _vp, _fp, _i64, _i, _i, _i, _i, _i, _vp],
    "stgcma_im2col": [_fp, _vp, _i, _i, _i, _i, _i, _i, _i64, _i, _vp],
    "stgcma_clip_embed_fwd": [_vp, _fp, _fp, _fp, _fp, _i64, _i, _i, _i, _i, _vp],
    "stgcma_clip_temb_bwd": [_fp, _fp, _i64, _i, _i, _i, _vp],
    "stgcma_cls_ln_fwd": [_fp, _fp, _fp, _fp, _fp, _fp, _i64, _i, _i, _i, _vp],
    "stgcma_cls_ln_bwd": [_fp, _fp, _fp, _fp, _fp, _fp, _vp, _fp, _fp, _i64, _i, _i, _i, _i, _vp],
}
EXPORTS = ["stgcma_last_error", "stgcma_version"] + list(_SIGNATURES)

_lib = None
_launches = 0  # number of kernel-launching ABI calls made so far (bench.py reports the per-step delta)


class StgcmaError(RuntimeError):
    pass


def load():
    """Load libstgcma.so (once).  Raises StgcmaError when it is absent -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise StgcmaError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C stg-cma_b200/csrc`).  The STG-CMA B200 path has no CPU / eager fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.stgcma_last_error.restype = ctypes.c_char_p
    lib.stgcma_last_error.argtypes = []
    lib.stgcma_version.restype = _i
    lib.stgcma_version.argtypes = []
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = _i
        fn.argtypes = argtypes
    _lib = lib
    return lib


def launches():
    return _launches


def add_launches(n):
    """CUDA-graph replays launch the captured kernels without passing through _call (engine.py)."""
    global _launches
    _launches += n


def profiling():
    return _prof is not None


def dt(t_or_dtype):
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    if d == torch.float32:
        return F32
    if d == torch.bfloat16:
        return BF16
    raise StgcmaError(f"unsupported activation dtype {d}")


def _p(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


_prof = None  # when a list: (name, start_event, end_event, flops, bytes) per ABI call


def profile_start():
    """Bracket every subsequent ABI call with CUDA events on the launching stream (bench.py uses this
    for the per-kernel time shares and the roofline of the dominant kernel)."""
    global _prof
    _prof = []


def profile_stop():
    """-> {kernel name: dict(ms, calls, flops, bytes)}; algorithmic flops / bytes as given by the wrappers."""
    global _prof
    rec, _prof = _prof, None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, flops, nbytes in rec or []:
        d = out.setdefault(name, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
        d["ms"] += e0.elapsed_time(e1)
        d["calls"] += 1
        d["flops"] += flops
        d["bytes"] += nbytes
    return out


def _call(name, *args, cost=(0.0, 0.0), tag=""):
    global _launches
    lib = load()
    stream = _stream()
    if _prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*args, stream)  # every kernel entry point takes the stream last
    if rc != 0:
        raise StgcmaError(f"{name} failed ({rc}): {lib.stgcma_last_error().decode()}")
    if _prof is not None:
        e1.record()
        _prof.append((name + tag, e0, e1, float(cost[0]), float(cost[1])))
    _launches += 1


def _es(t):
    return t.element_size()


_checked = set()


def check_device(dev_index):
    if dev_index in _checked:
        return
    lib = load()
    rc = lib.stgcma_check_device(dev_index)
    if rc != 0:
        raise StgcmaError(lib.stgcma_last_error().decode())
    _checked.add(dev_index)


def require_device(*tensors):
    """Inputs must be CUDA tensors on an sm_100 device: there is no CPU / eager fallback."""
    for t in tensors:
        if not t.is_cuda:
            raise StgcmaError("STG-CMA B200 path needs CUDA tensors: there is no CPU fallback")
    dev = tensors[0].device
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if any(t.device != dev for t in tensors):
        raise StgcmaError("STG-CMA B200 path: all inputs must live on the same device")
    if idx != torch.cuda.current_device():
        # kernels are enqueued on torch's CURRENT stream; pointers of another device would be launched in the wrong context
        raise StgcmaError(f"inputs are on cuda:{idx} but the current device is cuda:{torch.cuda.current_device()}: "
                          "call torch.cuda.set_device() (one process per GPU) before the step")
    check_device(idx)


# ---------------------------------------------------------------------------------------------
# thin typed wrappers (shapes are the caller's responsibility; see include/stgcma.h)
# ---------------------------------------------------------------------------------------------
def patch_embed_fwd(img, w, bias, gamma, beta, out, B, Cin, T, H, W, C):
    tok = B * T * (H // 4) * (W // 4)
    _call("stgcma_patch_embed_fwd", _p(img), _p(w), _p(bias), _p(gamma), _p(beta), _p(out), B, Cin, T, H, W, C,
          dt(out), cost=(2.0 * tok * Cin * 16 * C, 4.0 * B * Cin * T * H * W + tok * C * _es(out)))


def layernorm_fwd(x, gamma, beta, y, mean, rstd, rows, C, merge=0, R=0):
    _call("stgcma_layernorm_fwd", _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), rows, C, merge, R, dt(x), dt(y),
          cost=(0.0, rows * C * (_es(x) + _es(y))))


def layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx, rows, C, merge=0, R=0, dx_copy=None):
    _call("stgcma_layernorm_bwd", _p(dy), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dres), _p(dx), _p(dx_copy), rows, C,
          merge, R, dt(dy), dt(x), dt(dx_copy if dx_copy is not None else dy),
          cost=(0.0, rows * C * (_es(dy) + (2.0 + (dres is not None)) * _es(x) + (0 if dx_copy is None else _es(dx_copy)))))


def gemm(A, W, out, M, N, K, *, bias=None, res1=None, res2=None, rowscale=None, aux=None, out_pre=None, lda=None,
         ldw=None, ldo=None, groups=1, epi=EPI_LINEAR):
    """`out` (and `res1`, which must match it) may be fp32 while A / W are bf16: the residual stream."""
    if res1 is not None and res1.dtype != out.dtype:
        raise StgcmaError("gemm: res1 must have the dtype of out")
    _call("stgcma_gemm", _p(A), _p(W), _p(bias), _p(res1), _p(res2), _p(rowscale), _p(aux), _p(out), _p(out_pre), M, N,
          K, K if lda is None else lda, K if ldw is None else ldw, N if ldo is None else ldo, groups, epi, dt(A), dt(out),
          cost=(2.0 * M * N * K, _es(A) * (M * K + groups * N * K + M * N * (
              0.0 + (res2 is not None) + (aux is not None) + (out_pre is not None))) + _es(out) * M * N * (
              1.0 + (res1 is not None))), tag="[K<=192]" if K <= 192 else "[K>=256]")


def mlp_fused_ok(C, H, rp, dtype, backward=True):
    """Shapes the fused MLP kernels cover (include/stgcma.h); outside them the engine composes layernorm + gemm."""
    return (dtype == torch.bfloat16 and C % 32 == 0 and 32 <= C <= (256 if backward else 384) and H % 64 == 0
            and rp % 16 == 0 and 16 <= rp <= 128)


def mlp_fwd(x, gamma, beta, W1, b1, W2, b2, Wa, ba, y, hpre, h, mean, rstd, M, C, H, rp, groups=2):
    _call("stgcma_mlp_fwd", _p(x), _p(gamma), _p(beta), _p(W1), _p(b1), _p(W2), _p(b2), _p(Wa), _p(ba), _p(y), _p(hpre),
          _p(h), _p(mean), _p(rstd), M, C, H, rp, groups, dt(y),
          cost=(2.0 * M * (2 * C * H + C * rp), M * (4.0 * C + 2.0 * C + 4.0 * rp + 8.0)))


def mlp_bwd(dhpre, dxb, x, gamma, beta, mean, rstd, Wat, W1, b1, W2t, W1t, dxn, M, C, H, rp, groups=2):
    # algorithmic FLOPs: the two dX GEMMs of the MLP + the adapter dX GEMM (the fc1 recompute is not counted)
    _call("stgcma_mlp_bwd", _p(dhpre), _p(dxb), _p(x), _p(gamma), _p(beta), _p(mean), _p(rstd), _p(Wat), _p(W1), _p(b1),
          _p(W2t), _p(W1t), _p(dxn), M, C, H, rp, groups, dt(dxn),
          cost=(2.0 * M * (2 * C * H + C * rp), M * (4.0 * C + 2.0 * C + 2.0 * C + 2.0 * rp + 8.0)))


def gemm_tn(A, B, dW, M, Ka, Nb, *, lda, ldb, ldo, rowscale=None, csumA=None, csumB=None):
    _call("stgcma_gemm_tn", _p(A), _p(B), _p(rowscale), _p(dW), _p(csumA), _p(csumB), M, Ka, Nb, lda, ldb, ldo, dt(A),
          cost=(2.0 * M * Ka * Nb, _es(A) * M * (Ka + Nb)))


def window_attn_fwd(qkv, rpb, out, frames, R, ws, shift, heads):
    tok, C = frames * R * R, heads * 32
    _call("stgcma_window_attn_fwd", _p(qkv), _p(rpb), _p(out), frames, R, ws, shift, heads, dt(qkv),
          cost=(4.0 * tok * ws * ws * C, _es(qkv) * tok * 4 * C))


def window_attn_bwd(qkv, rpb, dout, dqkv, frames, R, ws, shift, heads):
    tok, C = frames * R * R, heads * 32
    _call("stgcma_window_attn_bwd", _p(qkv), _p(rpb), _p(dout), _p(dqkv), frames, R, ws, shift, heads, dt(qkv),
          cost=(8.0 * tok * ws * ws * C, _es(qkv) * tok * 7 * C))


def temporal_attn_fwd(qkv, tbl, out, G, B, T, L, heads):
    tok, C = G * B * T * L, heads * 32
    _call("stgcma_temporal_attn_fwd", _p(qkv), _p(tbl), _p(out), G, B, T, L, heads, dt(qkv),
          cost=(4.0 * tok * T * C, _es(qkv) * tok * 4 * C))


def temporal_attn_bwd(qkv, tbl, dout, dqkv, dtbl, G, B, T, L, heads):
    tok, C = G * B * T * L, heads * 32
    _call("stgcma_temporal_attn_bwd", _p(qkv), _p(tbl), _p(dout), _p(dqkv), _p(dtbl), G, B, T, L, heads, dt(qkv),
          cost=(8.0 * tok * T * C, _es(qkv) * tok * 7 * C))


# `r` = un-padded hidden width, used only for the algorithmic-FLOP accounting (8 N r per token pair:
# the reference's four bmm calls, Swin_AVE.py:753-757 / :801-805)
def xmodal_window_fwd(h, gate, out, attn_out, frames, R, ws, shift, rp, r=None):
    tok = frames * R * R
    _call("stgcma_xmodal_window_fwd", _p(h), _p(gate), _p(out), _p(attn_out), frames, R, ws, shift, rp, dt(h),
          cost=(8.0 * tok * ws * ws * (r or rp), _es(h) * 2 * tok * rp * 3))


def xmodal_window_bwd(h, hpre, gate, attn_out, dout, dh, dgate, frames, R, ws, shift, rp, r=None):
    tok = frames * R * R
    _call("stgcma_xmodal_window_bwd", _p(h), _p(hpre), _p(gate), _p(attn_out), _p(dout), _p(dh), _p(dgate), frames, R,
          ws, shift, rp, dt(h), cost=(16.0 * tok * ws * ws * (r or rp), _es(h) * 2 * tok * rp * 5))


def xmodal_global_fwd(h, gate, out, attn_out, lse, frames, L, rp, r=None, La=None):
    """La: rows per frame of the second modality when it differs from L (CLIP: 197 / 49); tensors are then the stacked
    blocks [frames * L, rp] + [frames * La, rp]."""
    La = L if La is None else La
    _call("stgcma_xmodal_global_fwd", _p(h), _p(gate), _p(out), _p(attn_out), _p(lse), frames, L, La, rp, dt(h),
          cost=(8.0 * frames * L * La * (r or rp), _es(h) * frames * (L + La) * rp * 3))


def xmodal_global_bwd(h, hpre, gate, attn_out, lse, dout, dh, dgate, work, frames, L, rp, r=None, La=None):
    La = L if La is None else La
    _call("stgcma_xmodal_global_bwd", _p(h), _p(hpre), _p(gate), _p(attn_out), _p(lse), _p(dout), _p(dh), _p(dgate),
          _p(work), frames, L, La, rp, dt(h), cost=(16.0 * frames * L * La * (r or rp), _es(h) * frames * (L + La) * rp * 5))


def head_pool_fwd(xn, feat, frames, N, C):
    _call("stgcma_head_pool_fwd", _p(xn), _p(feat), frames, N, C, dt(xn), cost=(0.0, _es(xn) * 2 * frames * (N + 1) * C))


def head_pool_bwd(dfeat, dxn, frames, N, C):
    _call("stgcma_head_pool_bwd", _p(dfeat), _p(dxn), frames, N, C, dt(dxn), cost=(0.0, _es(dxn) * 2 * frames * (N + 1) * C))


# ---- CLIP-ViT variant (MM_CLIP_AVE) ------------------------------------------------------------------------
def mha_fwd(qkv, out, lse, n_seq, seq_len, inner, heads, hd):
    """Attention core of nn.MultiheadAttention; `qkv` / `out` / `lse` may be row-offset views (one modality)."""
    tok = n_seq * seq_len
    _call("stgcma_mha_fwd", _p(qkv), _p(out), _p(lse), n_seq, seq_len, inner, heads, hd, dt(qkv),
          cost=(4.0 * tok * seq_len * heads * hd, _es(qkv) * tok * 4 * heads * hd))


def mha_bwd(qkv, out, lse, dout, dqkv, work, n_seq, seq_len, inner, heads, hd):
    tok = n_seq * seq_len
    _call("stgcma_mha_bwd", _p(qkv), _p(out), _p(lse), _p(dout), _p(dqkv), _p(work), n_seq, seq_len, inner, heads, hd,
          dt(qkv), cost=(10.0 * tok * seq_len * heads * hd, _es(qkv) * tok * 8 * heads * hd))


def im2col(img, out, B, Cin, T, H, W, P, ldo=None):
    K = Cin * P * P
    rows = B * T * ((H - P) // P + 1) * ((W - P) // P + 1)
    _call("stgcma_im2col", _p(img), _p(out), B, Cin, T, H, W, P, K if ldo is None else ldo, dt(out),
          cost=(0.0, rows * K * (4.0 + _es(out))))


def clip_embed_fwd(patches, cls, pos, temb, x0, BT, T, N, D):
    _call("stgcma_clip_embed_fwd", _p(patches), _p(cls), _p(pos), _p(temb), _p(x0), BT, T, N, D, dt(patches),
          cost=(0.0, BT * N * D * (4.0 + _es(patches))))


def clip_temb_bwd(dx0, dtemb, B, T, N, D):
    _call("stgcma_clip_temb_bwd", _p(dx0), _p(dtemb), B, T, N, D, cost=(0.0, 4.0 * B * T * N * D))


def cls_ln_fwd(x, gamma, beta, feat, mean, rstd, frames, Nv, Na, D):
    _call("stgcma_cls_ln_fwd", _p(x), _p(gamma), _p(beta), _p(feat), _p(mean), _p(rstd), frames, Nv, Na, D,
          cost=(0.0, 16.0 * frames * D))


def cls_ln_bwd(dfeat, x, mean, rstd, gamma, dx, dgamma, dbeta, frames, Nv, Na, D, dx_copy=None):
    _call("stgcma_cls_ln_bwd", _p(dfeat), _p(x), _p(mean), _p(rstd), _p(gamma), _p(dx), _p(dx_copy), _p(dgamma), _p(dbeta),
          frames, Nv, Na, D, dt(dx_copy if dx_copy is not None else dx), cost=(0.0, 24.0 * frames * D))


PACK_ENTRY = struct.Struct("<QQiiqqii")  # mirrors StgcmaPackEntry (48 bytes)


def pack_entry(src, dst, rows, cols, src_ld, dst_ld, transpose, dst_f32):
    return PACK_ENTRY.pack(src.data_ptr(), dst.data_ptr(), rows, cols, src_ld, dst_ld, int(transpose), int(dst_f32))


ADAM_ENTRY = struct.Struct("<QQQQqii")   # mirrors StgcmaAdamEntry (48 bytes)
ADAM_MAX_GROUPS = 8


def adam_entry(param, grad, exp_avg, exp_avg_sq, group):
    return ADAM_ENTRY.pack(param.data_ptr(), grad.data_ptr(), exp_avg.data_ptr(), exp_avg_sq.data_ptr(), param.numel(), group, 0)


def adam_step(table_dev, n, groups, grad_scale=1.0, numel=0):
    """groups: list (<= 8) of dicts lr / weight_decay / beta1 / beta2 / eps / bias_corr1 / bias_corr2_sqrt."""
    if len(groups) > ADAM_MAX_GROUPS:
        raise StgcmaError(f"adam_step: at most {ADAM_MAX_GROUPS} parameter groups")
    fields = ("lr", "weight_decay", "beta1", "beta2", "eps", "bias_corr1", "bias_corr2_sqrt", "one_minus_beta1", "one_minus_beta2")
    flat = []
    for f in fields:
        vals = [float(g[f]) if f in g else 1.0 - float(g["beta" + f[-1]]) for g in groups]
        flat += vals + [1.0] * (ADAM_MAX_GROUPS - len(vals))
    hp = (ctypes.c_float * len(flat))(*flat)      # StgcmaAdamGroups, passed by host pointer
    _call("stgcma_adam_step", _p(table_dev), n, ctypes.cast(hp, ctypes.c_void_p), ctypes.c_float(grad_scale),
          cost=(0.0, 28.0 * numel))


def adam_step_amp(table_dev, n, groups, steps, scale_dev, found_inf_dev, skipped_dev, numel=0):
    """Adam under GradScaler: `scale_dev` / `found_inf_dev` are the scaler's device scalars, `skipped_dev` an int32 device
    counter of skipped steps; `steps` = host step count per group slot (bias corrections are computed in the kernel)."""
    if len(groups) > ADAM_MAX_GROUPS:
        raise StgcmaError(f"adam_step: at most {ADAM_MAX_GROUPS} parameter groups")
    fields = ("lr", "weight_decay", "beta1", "beta2", "eps", "bias_corr1", "bias_corr2_sqrt", "one_minus_beta1", "one_minus_beta2")
    flat = []
    for f in fields:
        vals = [float(g[f]) if f in g else (1.0 - float(g["beta" + f[-1]]) if f.startswith("one_minus") else 1.0) for g in groups]
        flat += vals + [1.0] * (ADAM_MAX_GROUPS - len(vals))
    hp = (ctypes.c_float * len(flat))(*flat)
    st = (ctypes.c_int32 * ADAM_MAX_GROUPS)(*([int(t) for t in steps] + [1] * (ADAM_MAX_GROUPS - len(steps))))
    _call("stgcma_adam_step_amp", _p(table_dev), n, ctypes.cast(hp, ctypes.c_void_p), ctypes.cast(st, ctypes.c_void_p),
          _p(scale_dev), _p(found_inf_dev), _p(skipped_dev), cost=(0.0, 32.0 * numel))


def bucket_allreduce(peers_dev_ptr, world, out, n, scale, mc_ptr=None):
    """out[i] = scale * sum over the ranks' symmetric buckets (device array of `world` pointers at `peers_dev_ptr`, or the NVLS
    multicast address `mc_ptr`); the caller brackets the launch with cross-rank barriers."""
    require_device(out)
    _call("stgcma_bucket_allreduce", peers_dev_ptr, world, mc_ptr, _p(out), n, ctypes.c_float(scale),
          cost=(0.0, 4.0 * n * (world + 1)))


def bucket_allreduce_nvls(mc_in_ptr, mc_out_ptr, n, rank, world, scale):
    """Two-shot NVLS all-reduce: this rank reduces its 1 / world slice in the switch and multicasts it to every rank's output."""
    _call("stgcma_bucket_allreduce_nvls", mc_in_ptr, mc_out_ptr, n, rank, world, ctypes.c_float(scale), cost=(0.0, 8.0 * n / world))


def pack_params(table_dev, n, dtype_code):
    _call("stgcma_pack_params", _p(table_dev), n, dtype_code)
