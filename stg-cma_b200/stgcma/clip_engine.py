"""Forward / backward schedule of the CLIP-ViT AVE adapter train step over libstgcma (BASELINE.json configs[1]).

Host-side mirror of `MM_CLIP_AVE.forward(a, v, 'fusion'|'multimodal')` (reference AVE/model/CLIP_AVE.py:1089-1140 /
:1038-1087, block `ResidualAttentionBlock.forward` :359-429 / :256-289) and of the autograd graph torch builds for it
under the caller's freeze rule (AVE/traintest_adapt_ave29.py:38-61): dX flows through every frozen layer; dW exists
only for the adapters, the gates, `temporal_embedding[_audio]`, `ln_post` and the head.  Unlike the Swin model the
backward runs all the way to the token assembly, because the temporal embeddings are trainable.

Layout: activations are token-major `[rows, D]`; the rows of a modality are `(b, t, n)` (the reference's sequence-first
`[n, (b t), d]` is never materialised); the two modalities are STACKED along rows -- 197-token video block, then
49-token audio block -- so every frozen GEMM / LayerNorm runs once over `M = B*T*(Nv + Na)` rows.  The two blocks have
different row counts, so the per-modality adapter GEMMs and the attention cores are launched once per modality on row
slices.  The rearranges 'n (b t) d <-> t (b n) d' of the temporal branch are index arithmetic inside `stgcma_mha_*`
(`inner = N`).  All compute is in the C-ABI library; torch is used for memory, streams and the collective only.
"""
import math
import os

import torch

from . import lib as L_
from .engine import _Arena, _DpExchange, _StepGraphs, _rup, _EPS_PAD


class ClipAVEEngine:
    def __init__(self, module):
        self.m = module
        self.cfg = module.cfg
        self.arena_ = _Arena()
        self.frozen = None
        self.frozen_key = None
        self.train_pack = None
        self.train_key = None
        self.generation = 0
        self.dp = _DpExchange()       # mean of the gradient bucket over the ranks (peer-memory all-reduce; engine._DpExchange)
        self.side_stream_dw = os.environ.get("STGCMA_SIDE_DW", "1") != "0"
        self._side = {}
        self.graphs = _StepGraphs()   # CUDA-graph replay of the schedules for launch-bound steps (engine._StepGraphs)

    @property
    def arena(self):
        return self.arena_

    @property
    def dp_allreduce(self):
        return self.dp.enabled

    @dp_allreduce.setter
    def dp_allreduce(self, v):
        self.dp.enabled = bool(v)

    @property
    def graph_mode(self):
        return self.graphs.mode

    @graph_mode.setter
    def graph_mode(self, v):
        self.graphs.mode = v

    @property
    def _graphs(self):
        return self.graphs.entries

    @staticmethod
    def stream_dtype(adt):
        return torch.float32 if adt == torch.bfloat16 else adt

    # ------------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------------
    ADAPTERS = (("T", "T_Adapter", "T_Adapter_Audio"), ("S", "S_Adapter", "S_Adapter_Audio"),
                ("M", "MLP_Adapter", "MLP_Adapter_Audio"))

    def _blocks(self):
        return list(enumerate(self.m.transformer.resblocks))

    def _frozen_params(self):
        m = self.m
        ps = [m.class_embedding, m.positional_embedding, m.positional_embedding_audio, m.conv1.weight, m.conv1_audio.weight,
              m.ln_pre.weight, m.ln_pre.bias]
        for _, b in self._blocks():
            ps += [b.attn.in_proj_weight, b.attn.in_proj_bias, b.attn.out_proj.weight, b.attn.out_proj.bias, b.ln_1.weight,
                   b.ln_1.bias, b.ln_2.weight, b.ln_2.bias, b.mlp.c_fc.weight, b.mlp.c_fc.bias, b.mlp.c_proj.weight,
                   b.mlp.c_proj.bias]
        return ps

    def trainable_params(self):
        """Ordered list of (name, parameter) of everything the engine produces gradients for."""
        m = self.m
        out = [("temporal_embedding", m.temporal_embedding), ("temporal_embedding_audio", m.temporal_embedding_audio)]
        for i, b in self._blocks():
            p = f"transformer.resblocks.{i}."
            out += [(p + "gate_v", b.gate_v), (p + "gate_a", b.gate_a)]
            for _, nv, na in self.ADAPTERS:
                for nm in (nv, na):
                    ad = getattr(b, nm)
                    out += [(p + nm + ".D_fc1.weight", ad.D_fc1.weight), (p + nm + ".D_fc1.bias", ad.D_fc1.bias),
                            (p + nm + ".D_fc2.weight", ad.D_fc2.weight), (p + nm + ".D_fc2.bias", ad.D_fc2.bias)]
        out += [("ln_post.weight", m.ln_post.weight), ("ln_post.bias", m.ln_post.bias)]
        h = m.mlp_head
        out += [("mlp_head.0.weight", h[0].weight), ("mlp_head.0.bias", h[0].bias),
                ("mlp_head.2.weight", h[2].weight), ("mlp_head.2.bias", h[2].bias)]
        return out

    def _check_frozen(self):
        bad = [p for p in self._frozen_params() if p.requires_grad]
        if bad:
            raise L_.StgcmaError(
                "the B200 STG-CMA path implements the frozen-backbone adapter step (freeze_base=True, "
                "traintest_adapt_ave29.py:58-61): backbone parameters must have requires_grad=False; "
                f"{len(bad)} frozen-path tensors still require grad. Full fine-tuning is out of scope.")

    def _build_frozen(self, dev, adt):
        key = (str(dev), adt, tuple((p.data_ptr(), p._version) for p in self._frozen_params()))
        if self.frozen is not None and self.frozen_key == key:
            return self.frozen
        m, cfg, fz = self.m, self.cfg, {}

        def w(t):
            return t.detach().to(device=dev, dtype=adt).contiguous()

        def wt(t):
            return t.detach().to(device=dev, dtype=adt).t().contiguous()

        def f(t):
            return t.detach().to(device=dev, dtype=torch.float32).contiguous()

        def conv_w(t):   # [D, Cin, P, P] -> [D, Kp], K padded to a multiple of 8 (TMA row pitch)
            D = t.shape[0]
            K = t[0].numel()
            out = torch.zeros(D, _rup(K, 8), dtype=adt, device=dev)
            out[:, :K] = t.detach().reshape(D, K).to(device=dev, dtype=adt)
            return out

        fz["conv1"], fz["conv1_audio"] = conv_w(m.conv1.weight), conv_w(m.conv1_audio.weight)
        fz["cls"], fz["pos_v"], fz["pos_a"] = f(m.class_embedding), f(m.positional_embedding), f(m.positional_embedding_audio)
        fz["pre_g"], fz["pre_b"] = f(m.ln_pre.weight), f(m.ln_pre.bias)
        for i, b in self._blocks():
            fz[i] = dict(
                n1g=f(b.ln_1.weight), n1b=f(b.ln_1.bias), n2g=f(b.ln_2.weight), n2b=f(b.ln_2.bias),
                qkv_w=w(b.attn.in_proj_weight), qkv_wt=wt(b.attn.in_proj_weight), qkv_b=f(b.attn.in_proj_bias),
                proj_w=w(b.attn.out_proj.weight), proj_wt=wt(b.attn.out_proj.weight), proj_b=f(b.attn.out_proj.bias),
                fc1_w=w(b.mlp.c_fc.weight), fc1_wt=wt(b.mlp.c_fc.weight), fc1_b=f(b.mlp.c_fc.bias),
                fc2_w=w(b.mlp.c_proj.weight), fc2_wt=wt(b.mlp.c_proj.weight), fc2_b=f(b.mlp.c_proj.bias))
        self.frozen, self.frozen_key = fz, key
        return fz

    def _build_train_pack(self, dev, adt):
        tp = self.trainable_params()
        key = (str(dev), adt, tuple(p.data_ptr() for _, p in tp))
        if self.train_pack is not None and self.train_key == key:
            return self.train_pack
        for _, p in tp:
            if p.device != dev or p.dtype != torch.float32:
                raise L_.StgcmaError("trainable parameters must be fp32 tensors on the compute device")
        cfg, m = self.cfg, self.m
        D, r, T = cfg.embed_dim, cfg.r, cfg.num_frames
        rp = _rup(r, _EPS_PAD)
        entries, pk = [], dict(rp=rp, r=r)
        f32 = torch.float32

        def z(*shape, dtype=adt):
            return torch.zeros(*shape, dtype=dtype, device=dev)

        def add(src, dst, rows, cols, src_ld, dst_ld, transpose=False, f32=False):
            entries.append(L_.pack_entry(src, dst, rows, cols, src_ld, dst_ld, transpose, f32))

        temb = z(2, T, D, dtype=f32)
        add(m.temporal_embedding.data, temb[0], T, D, D, D, f32=True)
        add(m.temporal_embedding_audio.data, temb[1], T, D, D, D, f32=True)
        pk["temb"] = temb
        for i, b in self._blocks():
            d = {}
            gate = z(2, dtype=f32)
            add(b.gate_v.data, gate[0:1], 1, 1, 1, 1, f32=True)
            add(b.gate_a.data, gate[1:2], 1, 1, 1, 1, f32=True)
            d["gate"] = gate
            for tag, nv, na in self.ADAPTERS:
                w1, b1 = z(2, rp, D), z(2, rp, dtype=f32)
                w2, b2 = z(2, D, rp), z(2, D, dtype=f32)
                w1t, w2t = z(2, D, rp), z(2, rp, D)
                for g, nm in enumerate((nv, na)):
                    ad = getattr(b, nm)
                    add(ad.D_fc1.weight.data, w1[g], r, D, D, D)
                    add(ad.D_fc1.weight.data, w1t[g], r, D, D, rp, transpose=True)
                    add(ad.D_fc1.bias.data, b1[g], 1, r, r, rp, f32=True)
                    add(ad.D_fc2.weight.data, w2[g], D, r, r, rp)
                    add(ad.D_fc2.weight.data, w2t[g], D, r, r, D, transpose=True)
                    add(ad.D_fc2.bias.data, b2[g], 1, D, D, D, f32=True)
                d[tag] = dict(w1=w1, b1=b1, w2=w2, b2=b2, w1t=w1t, w2t=w2t)
            pk[i] = d
        post = z(2, D, dtype=f32)
        add(m.ln_post.weight.data, post[0], 1, D, D, D, f32=True)
        add(m.ln_post.bias.data, post[1], 1, D, D, D, f32=True)
        pk["post"] = post
        h0, h2 = m.mlp_head[0], m.mlp_head[2]
        nl = h2.weight.shape[0]
        lp = _rup(nl, 16)
        F2, Hd = h0.weight.shape[1], h0.weight.shape[0]
        hd = dict(lp=lp, nl=nl, w0=z(Hd, F2, dtype=f32), w0t=z(F2, Hd, dtype=f32), b0=z(Hd, dtype=f32),
                  w2=z(lp, Hd, dtype=f32), w2t=z(Hd, lp, dtype=f32), b2=z(lp, dtype=f32))
        add(h0.weight.data, hd["w0"], Hd, F2, F2, F2, f32=True)
        add(h0.weight.data, hd["w0t"], Hd, F2, F2, Hd, transpose=True, f32=True)
        add(h0.bias.data, hd["b0"], 1, Hd, Hd, Hd, f32=True)
        add(h2.weight.data, hd["w2"], nl, Hd, Hd, Hd, f32=True)
        add(h2.weight.data, hd["w2t"], nl, Hd, Hd, lp, transpose=True, f32=True)
        add(h2.bias.data, hd["b2"], 1, nl, nl, lp, f32=True)
        pk["head"] = hd
        raw = b"".join(entries)
        pk["_table"], pk["_n"] = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev), len(entries)
        offs, off = {}, 0
        for name, p in tp:
            offs[name] = (off, tuple(p.shape))
            off += p.numel()
        pk["_goffs"], pk["_gsize"] = offs, off
        self.train_pack, self.train_key = pk, key
        return pk

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def forward(self, a, v, mode, need_grad):
        cfg, m = self.cfg, self.m
        if mode not in ("fusion", "multimodal"):
            raise L_.StgcmaError(f"ftmode '{mode}' is outside the B200 hot path (fusion / multimodal only)")
        L_.require_device(a, v)
        dev, adt = v.device, m.compute_dtype
        if need_grad:
            self._check_frozen()
        fz = self._build_frozen(dev, adt)
        pk = self._build_train_pack(dev, adt)
        self.generation += 1
        B, Cin, T, H, W = v.shape
        Ha, Wa = cfg.audio_shape
        assert Cin == 3 and T == cfg.num_frames and H == cfg.input_resolution and W == cfg.input_resolution, \
            "video must be [B,3,T,H,W]"
        assert tuple(a.shape) == (B, T, Ha, Wa), f"audio must be [B,T,{Ha},{Wa}] (CLIP path)"
        a32 = a.detach().to(torch.float32).contiguous()
        v32 = v.detach().to(torch.float32).contiguous()

        def run(a_, v_):
            L_.pack_params(pk["_table"], pk["_n"], L_.dt(adt))
            return self._forward_part(a_, v_, mode, fz, pk)
        ent = None
        if self.graphs.wanted(B, dev):
            key = (mode, bool(need_grad), m.training, m.mlp_head[1].training, tuple(v32.shape), dev.index, adt, self.side_stream_dw)
            logits, part, ent = self.graphs.forward(key, fz, pk, a32, v32, run)
        else:
            logits, part = run(a32, v32)
        part["gen"], part["graph"] = self.generation, ent
        return logits, part

    def saved_masks(self, sv):
        """DropPath masks {layer: {'v': [Nv], 'a': [Na]}} (per TOKEN INDEX: the reference applies DropPath to the
        sequence-first tensor, CLIP_AVE.py:372,377) and the head-dropout keep mask (test hook)."""
        return dict(dp=sv["dp_tok"], keepmask=sv["keepmask"])

    def _forward_part(self, a32, v32, mode, fz, pk):
        cfg, m = self.cfg, self.m
        ar = self.arena_
        dev, adt = v32.device, m.compute_dtype
        sdt = self.stream_dtype(adt)
        f32 = torch.float32
        fusion = mode == "fusion"
        B, _, T, H, W = v32.shape
        Ha, Wa = cfg.audio_shape
        BT = B * T
        D, heads, hd, P = cfg.embed_dim, cfg.heads, cfg.head_dim, cfg.patch_size
        Nv, Na = cfg.nv, cfg.na
        Mv, Ma = BT * Nv, BT * Na
        M = Mv + Ma
        mods = ((0, Mv, Nv), (Mv, Ma, Na))          # (row offset, rows, tokens per frame) of video / audio
        rp, r = pk["rp"], pk["r"]
        training = m.training
        sv = dict(B=B, mode=mode, dev=dev, adt=adt, sdt=sdt, dp={}, dp_tok={})

        def buf(name, *shape, dtype=adt, zero=False):
            return ar.get(name, shape, dtype, dev, zero)

        def per_mod(fn):
            for g, (off, rows, n) in enumerate(mods):
                fn(g, off, rows, n)

        # ---- patchify (conv1 / conv1_audio as im2col + GEMM), token assembly, ln_pre ----
        x0 = buf("x0", M, D, dtype=f32)
        for g, (img, cin, hh, ww, wname, pos) in enumerate(((v32, 3, H, W, "conv1", "pos_v"), (a32, 1, Ha, Wa, "conv1_audio", "pos_a"))):
            off, rows, n = mods[g]
            Kp = fz[wname].shape[1]
            col = buf(f"col{g}", BT * (n - 1), Kp, zero=True)
            L_.im2col(img, col, B, cin, T, hh, ww, P, ldo=Kp)
            pat = buf(f"pat{g}", BT * (n - 1), D)
            L_.gemm(col, fz[wname], pat, BT * (n - 1), D, Kp)
            L_.clip_embed_fwd(pat, fz["cls"], fz[pos], pk["temb"][g], x0[off:off + rows], BT, T, n, D)
        x = buf("x_in", M, D, dtype=sdt)
        L_.layernorm_fwd(x0, fz["pre_g"], fz["pre_b"], x, buf("pre_mean", M, dtype=f32), buf("pre_rstd", M, dtype=f32), M, D)

        for i, blk in self._blocks():
            f, p = fz[i], pk[i]
            k = f"b{i}_"
            # ---------------- (T) temporal attention + T_Adapter (no skip) + DropPath ----------------
            xn = buf("xn", M, D)
            L_.layernorm_fwd(x, f["n1g"], f["n1b"], xn, buf(k + "t_mean", M, dtype=f32), buf(k + "t_rstd", M, dtype=f32), M, D)
            qkv_t = buf(k + "qkv_t", M, 3 * D)
            L_.gemm(xn, f["qkv_w"], qkv_t, M, 3 * D, D, bias=f["qkv_b"])
            o_t, lse_t = buf(k + "o_t", M, D), buf(k + "lse_t", M, heads, dtype=f32)
            per_mod(lambda g, off, rows, n: L_.mha_fwd(qkv_t[off:off + rows], o_t[off:off + rows], lse_t[off:off + rows],
                                                      B * n, T, n, heads, hd))
            rt = buf(k + "rt", M, D)
            L_.gemm(o_t, f["proj_w"], rt, M, D, D, bias=f["proj_b"])
            hpre, h = buf(k + "t_hpre", M, rp), buf(k + "t_h", M, rp)
            A = p["T"]
            per_mod(lambda g, off, rows, n: L_.gemm(rt[off:off + rows], A["w1"][g], h[off:off + rows], rows, rp, D, bias=A["b1"][g],
                                                   out_pre=hpre[off:off + rows], epi=L_.EPI_GELU))
            dp = None
            if training and blk.drop_path_rate > 0.0:
                keep = 1.0 - blk.drop_path_rate
                toks = [torch.empty(n, device=dev, dtype=f32).bernoulli_(keep).div_(keep) for _, _, n in mods]
                dp = torch.cat([tk.view(1, n).expand(BT, n).reshape(-1) for tk, (_, _, n) in zip(toks, mods)])
                sv["dp_tok"][i] = dict(v=toks[0], a=toks[1])
            sv["dp"][i] = dp
            x1 = buf(k + "x1", M, D, dtype=sdt)
            per_mod(lambda g, off, rows, n: L_.gemm(h[off:off + rows], A["w2"][g], x1[off:off + rows], rows, D, rp, bias=A["b2"][g],
                                                   res1=x[off:off + rows], rowscale=None if dp is None else dp[off:off + rows]))
            sv[k + "x0"] = x
            x = x1
            # ---------------- (S) spatial attention + S_Adapter with cross-modal attention ----------------
            xn = buf("xn", M, D)
            L_.layernorm_fwd(x, f["n1g"], f["n1b"], xn, buf(k + "s_mean", M, dtype=f32), buf(k + "s_rstd", M, dtype=f32), M, D)
            qkv = buf(k + "qkv", M, 3 * D)
            L_.gemm(xn, f["qkv_w"], qkv, M, 3 * D, D, bias=f["qkv_b"])
            o_s, lse_s = buf(k + "o_s", M, D), buf(k + "lse_s", M, heads, dtype=f32)
            per_mod(lambda g, off, rows, n: L_.mha_fwd(qkv[off:off + rows], o_s[off:off + rows], lse_s[off:off + rows],
                                                      BT, n, 1, heads, hd))
            vs = buf(k + "vs", M, D)
            L_.gemm(o_s, f["proj_w"], vs, M, D, D, bias=f["proj_b"])
            hpre, h = buf(k + "s_hpre", M, rp), buf(k + "s_h", M, rp)
            A = p["S"]
            per_mod(lambda g, off, rows, n: L_.gemm(vs[off:off + rows], A["w1"][g], h[off:off + rows], rows, rp, D, bias=A["b1"][g],
                                                   out_pre=hpre[off:off + rows], epi=L_.EPI_GELU))
            if fusion:
                h2, att = buf(k + "s_h2", M, rp), buf(k + "s_att", M, rp)
                L_.xmodal_global_fwd(h, p["gate"], h2, att, buf(k + "s_lse", M, dtype=f32), BT, Nv, rp, r=r, La=Na)
            else:
                h2 = h
            x2 = buf(k + "x2", M, D, dtype=sdt)
            per_mod(lambda g, off, rows, n: L_.gemm(h2[off:off + rows], A["w2"][g], x2[off:off + rows], rows, D, rp, bias=A["b2"][g],
                                                   res1=x[off:off + rows], res2=vs[off:off + rows]))
            sv[k + "x1"] = x
            x = x2
            # ---------------- (G) frozen MLP (QuickGELU) + MLP_Adapter with cross-modal attention ----------------
            xn = buf("xn", M, D)
            L_.layernorm_fwd(x, f["n2g"], f["n2b"], xn, buf(k + "g_mean", M, dtype=f32), buf(k + "g_rstd", M, dtype=f32), M, D)
            hid = f["fc1_w"].shape[0]
            # fpre holds QuickGELU'(c_fc(xn)): the forward epilogue has sigmoid(1.702 x) at hand, so the backward of the
            # activation is a plain multiply instead of a second exp + reciprocal per element (DQGELU ran at 699 TFLOP/s)
            fpre, fact = buf(k + "fpre", M, hid), buf("fact", M, hid)
            L_.gemm(xn, f["fc1_w"], fact, M, hid, D, bias=f["fc1_b"], out_pre=fpre, epi=L_.EPI_QGELU_G)
            vn = buf(k + "vn", M, D)
            L_.gemm(fact, f["fc2_w"], vn, M, D, hid, bias=f["fc2_b"])
            hpre, h = buf(k + "g_hpre", M, rp), buf(k + "g_h", M, rp)
            A = p["M"]
            per_mod(lambda g, off, rows, n: L_.gemm(vn[off:off + rows], A["w1"][g], h[off:off + rows], rows, rp, D, bias=A["b1"][g],
                                                   out_pre=hpre[off:off + rows], epi=L_.EPI_GELU))
            if fusion:
                h2, att = buf(k + "g_h2", M, rp), buf(k + "g_att", M, rp)
                L_.xmodal_global_fwd(h, p["gate"], h2, att, buf(k + "g_lse", M, dtype=f32), BT, Nv, rp, r=r, La=Na)
            else:
                h2 = h
            x3 = buf(k + "x3", M, D, dtype=sdt)
            per_mod(lambda g, off, rows, n: L_.gemm(h2[off:off + rows], A["w2"][g], x3[off:off + rows], rows, D, rp, bias=A["b2"][g],
                                                   res1=x[off:off + rows], res2=vn[off:off + rows]))
            sv[k + "x2"] = x
            x = x3

        # ---------------- head: ln_post on the class tokens, cat(a, v), MLP ----------------
        hdp = pk["head"]
        feat = buf("h_feat", BT, 2 * D, dtype=f32)
        L_.cls_ln_fwd(x, pk["post"][0], pk["post"][1], feat, buf("h_mean", 2 * BT, dtype=f32), buf("h_rstd", 2 * BT, dtype=f32),
                      BT, Nv, Na, D)
        Hd = hdp["w0"].shape[0]
        hh = buf("h_hidden", BT, Hd, dtype=f32)
        drop = m.mlp_head[1]
        keepmask = None
        if drop.training and drop.p > 0.0:
            keepmask = torch.empty(BT, Hd, device=dev, dtype=f32).bernoulli_(1.0 - drop.p).div_(1.0 - drop.p)
            L_.gemm(feat, hdp["w0"], hh, BT, Hd, 2 * D, bias=hdp["b0"], aux=keepmask, epi=L_.EPI_MULAUX)
        else:
            L_.gemm(feat, hdp["w0"], hh, BT, Hd, 2 * D, bias=hdp["b0"])
        sv["keepmask"] = keepmask
        lp, nl = hdp["lp"], hdp["nl"]
        logits_p = buf("h_logits", BT, lp, dtype=f32)
        L_.gemm(hh, hdp["w2"], logits_p, BT, lp, Hd, bias=hdp["b2"])
        sv["x_final"] = x
        return logits_p[:, :nl].clone(), sv

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def backward(self, sv, dlogits):
        if sv["gen"] != self.generation:
            raise L_.StgcmaError("backward() of a stale forward: the activation arena holds one step at a time "
                                 "(run forward -> backward -> forward ...)")
        dev = sv["dev"]
        pk = self.train_pack
        gbuf = self.dp.bucket(pk["_gsize"], dev)
        if gbuf is None:
            gbuf = self.arena_.get("grad_bucket", (pk["_gsize"],), torch.float32, dev)
        out = None
        if sv.get("graph") is not None:
            out = self.graphs.backward(sv["graph"], dlogits, gbuf, self._backward_part)
        else:
            gbuf.zero_()
            self._backward_part(sv, dlogits, gbuf)
        if self.dp.active():
            out = self.dp.reduce(gbuf)
        elif out is None:
            out = gbuf.clone()
        goffs = pk["_goffs"]
        grads = []
        for name, prm in self.trainable_params():
            off, shape = goffs[name]
            grads.append(out[off:off + math.prod(shape)].view(shape) if prm.requires_grad else None)
        return grads

    def _backward_part(self, sv, dlogits, gbuf):
        cfg, m = self.cfg, self.m
        ar = self.arena_
        dev, adt, sdt, B = sv["dev"], sv["adt"], sv["sdt"], sv["B"]
        fz, pk = self.frozen, self.train_pack
        fusion = sv["mode"] == "fusion"
        f32 = torch.float32
        T = cfg.num_frames
        BT = B * T
        D, heads, hd = cfg.embed_dim, cfg.heads, cfg.head_dim
        Nv, Na = cfg.nv, cfg.na
        Mv, Ma = BT * Nv, BT * Na
        M = Mv + Ma
        mods = ((0, Mv, Nv), (Mv, Ma, Na))
        rp, r = pk["rp"], pk["r"]

        def buf(name, *shape, dtype=adt, zero=False):
            return ar.get(name, shape, dtype, dev, zero)

        def slot(name):
            return ar.slots[name]

        def per_mod(fn):
            for g, (off, rows, n) in enumerate(mods):
                fn(g, off, rows, n)

        goffs = pk["_goffs"]

        def g_(name):
            off, shape = goffs[name]
            return gbuf[off:off + math.prod(shape)]

        # ---------------- head ----------------
        hdp = pk["head"]
        lp, nl = hdp["lp"], hdp["nl"]
        Hd = hdp["w0"].shape[0]
        dlog = buf("d_logits", BT, lp, dtype=f32, zero=True)
        dlog[:, :nl].copy_(dlogits)
        hh, feat = slot("h_hidden"), slot("h_feat")
        L_.gemm_tn(dlog, hh, g_("mlp_head.2.weight"), BT, nl, Hd, lda=lp, ldb=Hd, ldo=Hd, csumA=g_("mlp_head.2.bias"))
        dh = buf("d_hidden", BT, Hd, dtype=f32)
        if sv["keepmask"] is not None:
            L_.gemm(dlog, hdp["w2t"], dh, BT, Hd, lp, aux=sv["keepmask"], epi=L_.EPI_MULAUX)
        else:
            L_.gemm(dlog, hdp["w2t"], dh, BT, Hd, lp)
        L_.gemm_tn(dh, feat, g_("mlp_head.0.weight"), BT, Hd, 2 * D, lda=Hd, ldb=2 * D, ldo=2 * D, csumA=g_("mlp_head.0.bias"))
        dfeat = buf("d_feat", BT, 2 * D, dtype=f32)
        L_.gemm(dh, hdp["w0t"], dfeat, BT, 2 * D, Hd)
        # dx = gradient of the residual stream (fp32); dxb = its activation-dtype copy (the GEMM operand).  Only the class
        # tokens carry a gradient at this point.  The three sub-blocks rotate over three buffer pairs (dD -> dE -> dB -> dD)
        # so that a side-stream dW kernel never reads a buffer the main chain rewrites before the next join.
        dx, dxb = buf("dD", M, D, dtype=sdt), buf("dDb", M, D)
        dx.zero_()
        dxb.zero_()
        L_.cls_ln_bwd(dfeat, sv["x_final"], slot("h_mean"), slot("h_rstd"), pk["post"][0], dx, g_("ln_post.weight"),
                      g_("ln_post.bias"), BT, Nv, Na, D, dx_copy=dxb)

        side = None
        main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None
        if self.side_stream_dw and main is not None:
            key = main.cuda_stream
            if key not in self._side:
                self._side[key] = torch.cuda.Stream(device=dev)
            side = self._side[key]
        pending = []

        def join_side():
            while pending:
                main.wait_event(pending.pop())

        def adapter_grads(names, d_out, h_in, d_hpre, x_in, rowscale):
            """dW2/db2 from (d_out, h_in); dW1/db1 from (d_hpre, x_in), per modality (dW is off the dX critical path)."""
            def run():
                for g, (off, rows, n) in enumerate(mods):
                    nm = names[g]
                    rs = None if rowscale is None else rowscale[off:off + rows]
                    L_.gemm_tn(d_out[off:off + rows], h_in[off:off + rows], g_(nm + ".D_fc2.weight"), rows, D, r, lda=D, ldb=rp,
                               ldo=r, rowscale=rs, csumA=g_(nm + ".D_fc2.bias"))
                    L_.gemm_tn(d_hpre[off:off + rows], x_in[off:off + rows], g_(nm + ".D_fc1.weight"), rows, r, D, lda=rp, ldb=D,
                               ldo=D, csumA=g_(nm + ".D_fc1.bias"))
            if side is None:
                run()
                return
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                run()
                done = torch.cuda.Event()
                done.record(side)
            pending.append(done)

        work = buf("mha_work", M, heads, dtype=f32)
        xwork = buf("xwork", M, dtype=f32)
        for i, blk in reversed(self._blocks()):
            f, p = fz[i], pk[i]
            k = f"b{i}_"
            pfx = f"transformer.resblocks.{i}."
            hid = f["fc1_w"].shape[0]
            assert goffs[pfx + "gate_a"][0] == goffs[pfx + "gate_v"][0] + 1
            dgate2 = gbuf[goffs[pfx + "gate_v"][0]:goffs[pfx + "gate_v"][0] + 2]
            dh2, dhpre = buf("dh", M, rp), buf("dhpre", M, rp)

            def adapter_in_bwd(A, dsrc, tag):
                """gradient wrt the adapter hidden pre-activation from the gradient of the adapter output (dsrc)."""
                if fusion:
                    per_mod(lambda g, off, rows, n: L_.gemm(dsrc[off:off + rows], A["w2t"][g], dh2[off:off + rows], rows, rp, D))
                    L_.xmodal_global_bwd(slot(k + tag + "_h"), slot(k + tag + "_hpre"), p["gate"], slot(k + tag + "_att"),
                                         slot(k + tag + "_lse"), dh2, dhpre, dgate2, xwork, BT, Nv, rp, r=r, La=Na)
                    return slot(k + tag + "_h2")
                hp = slot(k + tag + "_hpre")
                per_mod(lambda g, off, rows, n: L_.gemm(dsrc[off:off + rows], A["w2t"][g], dhpre[off:off + rows], rows, rp, D,
                                                       aux=hp[off:off + rows], epi=L_.EPI_DGELU))
                return slot(k + tag + "_h")

            # ---------------- (G) backward ----------------
            join_side()
            A = p["M"]
            h2 = adapter_in_bwd(A, dxb, "g")
            adapter_grads((pfx + "MLP_Adapter", pfx + "MLP_Adapter_Audio"), dxb, h2, dhpre, slot(k + "vn"), None)
            dy = buf("dA", M, D)
            per_mod(lambda g, off, rows, n: L_.gemm(dhpre[off:off + rows], A["w1t"][g], dy[off:off + rows], rows, D, rp,
                                                   res1=dxb[off:off + rows]))
            dfpre = buf("dfpre", M, hid)
            L_.gemm(dy, f["fc2_wt"], dfpre, M, hid, D, aux=slot(k + "fpre"), epi=L_.EPI_MULAUX)
            dxn = buf("dC", M, D)
            L_.gemm(dfpre, f["fc1_wt"], dxn, M, D, hid)
            dx2, dx2b = buf("dE", M, D, dtype=sdt), buf("dEb", M, D)
            L_.layernorm_bwd(dxn, sv[k + "x2"], slot(k + "g_mean"), slot(k + "g_rstd"), f["n2g"], dx, dx2, M, D, dx_copy=dx2b)
            # ---------------- (S) backward ----------------
            join_side()
            A = p["S"]
            h2 = adapter_in_bwd(A, dx2b, "s")
            adapter_grads((pfx + "S_Adapter", pfx + "S_Adapter_Audio"), dx2b, h2, dhpre, slot(k + "vs"), None)
            dvs = buf("dA", M, D)
            per_mod(lambda g, off, rows, n: L_.gemm(dhpre[off:off + rows], A["w1t"][g], dvs[off:off + rows], rows, D, rp,
                                                   res1=dx2b[off:off + rows]))
            do = buf("dC", M, D)
            L_.gemm(dvs, f["proj_wt"], do, M, D, D)
            dqkv = buf("dqkv", M, 3 * D)
            qkv, o_s, lse_s = slot(k + "qkv"), slot(k + "o_s"), slot(k + "lse_s")
            per_mod(lambda g, off, rows, n: L_.mha_bwd(qkv[off:off + rows], o_s[off:off + rows], lse_s[off:off + rows],
                                                      do[off:off + rows], dqkv[off:off + rows], work[off:off + rows],
                                                      BT, n, 1, heads, hd))
            dxn = buf("dA", M, D)
            L_.gemm(dqkv, f["qkv_wt"], dxn, M, D, 3 * D)
            dx1, dx1b = buf("dB", M, D, dtype=sdt), buf("dBb", M, D)
            L_.layernorm_bwd(dxn, sv[k + "x1"], slot(k + "s_mean"), slot(k + "s_rstd"), f["n1g"], dx2, dx1, M, D, dx_copy=dx1b)
            # ---------------- (T) backward ----------------
            join_side()
            A = p["T"]
            dp = sv["dp"][i]
            hp = slot(k + "t_hpre")
            per_mod(lambda g, off, rows, n: L_.gemm(dx1b[off:off + rows], A["w2t"][g], dhpre[off:off + rows], rows, rp, D,
                                                   rowscale=None if dp is None else dp[off:off + rows], aux=hp[off:off + rows],
                                                   epi=L_.EPI_DGELU))
            adapter_grads((pfx + "T_Adapter", pfx + "T_Adapter_Audio"), dx1b, slot(k + "t_h"), dhpre, slot(k + "rt"), dp)
            drt = buf("dA", M, D)
            per_mod(lambda g, off, rows, n: L_.gemm(dhpre[off:off + rows], A["w1t"][g], drt[off:off + rows], rows, D, rp))
            do = buf("dC", M, D)
            L_.gemm(drt, f["proj_wt"], do, M, D, D)
            dqkv = buf("dqkv", M, 3 * D)
            qkv_t, o_t, lse_t = slot(k + "qkv_t"), slot(k + "o_t"), slot(k + "lse_t")
            per_mod(lambda g, off, rows, n: L_.mha_bwd(qkv_t[off:off + rows], o_t[off:off + rows], lse_t[off:off + rows],
                                                      do[off:off + rows], dqkv[off:off + rows], work[off:off + rows],
                                                      B * n, T, n, heads, hd))
            dxn = buf("dA", M, D)
            L_.gemm(dqkv, f["qkv_wt"], dxn, M, D, 3 * D)
            dx0, dx0b = buf("dD", M, D, dtype=sdt), buf("dDb", M, D)
            L_.layernorm_bwd(dxn, sv[k + "x0"], slot(k + "t_mean"), slot(k + "t_rstd"), f["n1g"], dx1, dx0, M, D, dx_copy=dx0b)
            dx, dxb = dx0, dx0b

        # ---------------- ln_pre backward -> gradients of the temporal embeddings ----------------
        join_side()
        dxe = buf("dE", M, D, dtype=sdt)
        L_.layernorm_bwd(dx, slot("x0"), slot("pre_mean"), slot("pre_rstd"), fz["pre_g"], None, dxe, M, D)
        L_.clip_temb_bwd(dxe[:Mv], g_("temporal_embedding"), B, T, Nv, D)
        L_.clip_temb_bwd(dxe[Mv:], g_("temporal_embedding_audio"), B, T, Na, D)
