"""Forward / backward schedule of the Swin AVE adapter train step over libstgcma.

Host-side mirror of `SwinTransformer2D_Adapter_New.forward(a, v, 'fusion'|'multimodal')`
(reference AVE/model/Swin_AVE.py:1565-1599, block :693-813 / :490-590) and of the autograd graph
torch would build for it under the caller's freeze rule (AVE/traintest_adapt_ave29.py:38-61):
dX flows through every frozen layer, dW exists only for adapters, gates, temporal bias tables
and the head.

Design points (B200-first, not a translation of the reference's module tree):
  * both modalities are stacked along the row axis ([video rows | audio rows]) so every frozen
    GEMM / LayerNorm / attention runs once over 2x the rows; the per-modality adapter weights
    are the two "groups" of one grouped GEMM.
  * roll / window_partition / window_reverse / rearrange never move data: activations stay
    token-major and the attention kernels gather / scatter by index.
  * all activations live in a persistent arena (fixed addresses step to step: allocator-free,
    CUDA-graph friendly); saved-for-backward tensors are arena slots too.
  * trainable parameters are re-packed (cast / pad / transpose) once per step by ONE kernel from a
    device-resident table; gradients land in ONE flat fp32 bucket that is all-reduced once when
    torch.distributed is initialised (data parallel over clips).
All compute is in the C-ABI library; torch is used for memory, streams and the collective only.
"""
import math
import os

import torch

from . import lib as L_

_EPS_PAD = 16  # adapter hidden width is padded to a multiple of 16 (UMMA K/N granularity)


def _rup(x, m):
    return (x + m - 1) // m * m


_POISON = os.environ.get("STGCMA_POISON", "0") == "1"


class _Arena:
    """Named persistent device buffers; a slot is re-allocated only when its shape/dtype changes."""

    def __init__(self):
        self.slots = {}

    def get(self, name, shape, dtype, device, zero=False):
        t = self.slots.get(name)
        shape = tuple(int(s) for s in shape)
        if t is None or tuple(t.shape) != shape or t.dtype != dtype or t.device != device:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            if _POISON and not zero and t.is_floating_point():
                t.fill_(float("nan"))   # debug: a kernel that reads a buffer before anything wrote it shows up as NaN
            self.slots[name] = t
        return t

    def bytes(self):
        return sum(t.numel() * t.element_size() for t in self.slots.values())

    def clear(self):
        self.slots.clear()


class _DpExchange:
    """The one exchange step of the data-parallel path: the MEAN of the flat gradient bucket over the ranks.

    On a CUDA box the bucket lives in symmetric (peer-mapped) memory (`torch.distributed._symmetric_memory`): the weight-gradient
    kernels accumulate straight into it, and `stgcma_bucket_allreduce` -- one launch between two cross-rank barriers -- reads the
    W buckets over NVLink / NVSwitch and writes the mean into a local tensor (8.9 MB for Swin-T: a latency-bound one-shot
    all-reduce, no NCCL call, no staging copy; 38 us against NCCL's 58 us on 2 GPUs).  From 4 ranks on, reading W whole buckets
    per rank loses to a reduce-scatter / all-gather shape (4 GPUs: 74 us one-shot, 64 us two-shot, NCCL 93 us; 8 GPUs: 177 / 54 /
    85 us), so the default ("auto") switches to the two-shot NVLS kernel when
    the box has NVSwitch multicast: each rank reduces 1 / W of the bucket IN THE SWITCH (multimem.ld_reduce) and multicasts the
    result into every rank's output copy (multimem.st).  STGCMA_DP=peer | nvls1 (one-shot through the switch) | nvls | nccl
    force a mode; `torch.distributed.all_reduce` is also what runs when symmetric memory cannot be set up (logged once) and on
    the CPU (gloo) schedule tests."""

    def __init__(self):
        self.mode = os.environ.get("STGCMA_DP", "auto")
        self.enabled = True          # engine.dp_allreduce
        self.t = self.full = self.hdl = self.osym = self.ohdl = None
        self.failed = False

    @staticmethod
    def world():
        d = torch.distributed
        return d.get_world_size() if d.is_available() and d.is_initialized() else 1

    def active(self):
        return self.enabled and self.world() > 1

    def bucket(self, n, dev):
        """The symmetric gradient bucket ([n] fp32) or None (use a local buffer + NCCL / gloo)."""
        if not self.active() or dev.type != "cuda" or self.mode == "nccl" or self.failed:
            return None
        if self.t is None or self.t.numel() != n or self.t.device != dev:
            try:
                import torch.distributed._symmetric_memory as symm
                q = 4 * self.world()                                   # whole float4s, equal slices per rank
                full = symm.empty((n + q - 1) // q * q, dtype=torch.float32, device=dev)
                full.zero_()
                self.hdl = symm.rendezvous(full, torch.distributed.group.WORLD)        # collective: every rank is here
                self.full, self.t = full, full[:n]
                self.osym = self.ohdl = None
                if self.hdl.multicast_ptr and self.mode in ("auto", "nvls"):
                    self.osym = symm.empty(full.numel(), dtype=torch.float32, device=dev)   # the two-shot kernel's output copy
                    self.ohdl = symm.rendezvous(self.osym, torch.distributed.group.WORLD)
            except Exception as e:  # noqa: BLE001 -- no peer mapping on this box: the NCCL collective does the same job
                self.failed, self.t, self.full, self.hdl = True, None, None, None
                print(f"[stgcma] symmetric memory unavailable ({type(e).__name__}: {e}); gradient bucket goes through NCCL")
                return None
        return self.t

    def reduce(self, gbuf):
        """-> a fresh tensor holding the mean of `gbuf` over the ranks."""
        W = self.world()
        if self.hdl is not None and gbuf is self.t:
            npad = self.full.numel()
            two_shot = self.ohdl is not None and (self.mode == "nvls" or (self.mode == "auto" and W >= 4))
            self.hdl.barrier(channel=0)       # every rank's bucket is complete
            if two_shot:
                L_.bucket_allreduce_nvls(self.hdl.multicast_ptr, self.ohdl.multicast_ptr, npad, self.hdl.rank, W, 1.0 / W)
                self.hdl.barrier(channel=1)   # every slice has landed in every copy; the buckets may be zeroed again
                return self.osym[:gbuf.numel()].clone()
            out = torch.empty(npad, dtype=torch.float32, device=gbuf.device)
            mc = (self.hdl.multicast_ptr or None) if self.mode == "nvls1" else None   # 0: no NVLS multicast mapping on this box
            L_.bucket_allreduce(self.hdl.buffer_ptrs_dev, W, out, npad, 1.0 / W, mc_ptr=mc)
            self.hdl.barrier(channel=1)       # ... and may be zeroed again only after every rank has read it
            return out[:gbuf.numel()]
        out = gbuf.clone()
        torch.distributed.all_reduce(out)
        out.div_(W)
        return out


class _StepGraphs:
    """CUDA-graph capture / replay of an engine's forward and backward schedules (shared by the Swin and CLIP engines).

    A step is 500 - 800 launches plus the torch ops around them; below ~8 clips per GPU the host cannot enqueue them as
    fast as the GPU retires them (BASELINE configs[0] runs 2 clips; the authors trained with 1 - 3).  Each schedule is
    captured once per (mode, shape, flags) key after two eager warm-up steps (arena allocation, kernel attributes, lazy
    module loading) and replayed: every buffer lives in the persistent arena, so the captured addresses stay valid; the
    DropPath / dropout masks are drawn inside the captured schedule by torch's graph-safe generator.
    STGCMA_GRAPH=1: always, 0: never, default: steps of <= 8 clips."""

    def __init__(self):
        self.mode = os.environ.get("STGCMA_GRAPH", "auto")
        self.max_clips = 8
        self.max_entries = 6         # captured (mode, shape, flags) keys kept alive; each holds its graphs' private pool
        self.entries = {}

    def wanted(self, B, dev):
        if dev.type != "cuda" or self.mode == "0" or L_.profiling():
            return False
        return self.mode == "1" or B <= self.max_clips

    def forward(self, key, fz, pk, a32, v32, run):
        """run(a, v) -> (logits, part) launches the forward schedule.  Returns (logits, part, entry or None)."""
        ent = self.entries.pop(key, None)
        if ent is None or ent["fz"] is not fz or ent["pk"] is not pk:
            # (new weight caches invalidate the captured addresses: start over)
            ent = dict(fz=fz, pk=pk, seen=0, fwd=None, bwd=None)
        self.entries[key] = ent                 # most recently used last
        while len(self.entries) > self.max_entries:
            self.entries.pop(next(iter(self.entries)))   # e.g. the odd last batch of every epoch: drop the oldest capture
        if ent["seen"] < 2:
            ent["seen"] += 1
            logits, part = run(a32, v32)
            return logits, part, None
        if ent["fwd"] is None:
            ent["a"], ent["v"] = a32.clone(), v32.clone()       # static inputs of the captured schedule
            torch.cuda.current_stream(v32.device).synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L_.launches()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                logits, part = run(ent["a"], ent["v"])
            ent.update(fwd=g, logits=logits, part=part, fwd_launches=L_.launches() - n0)
            L_.add_launches(-ent["fwd_launches"])               # capture launches nothing; the replay below is counted
        else:
            ent["a"].copy_(a32)
            ent["v"].copy_(v32)
        ent["fwd"].replay()
        L_.add_launches(ent["fwd_launches"])
        return ent["logits"].clone(), ent["part"], ent

    def backward(self, ent, dlogits, gbuf, run):
        """run(part, dlogits, gbuf) launches the backward schedule into the (zeroed) bucket.  -> a fresh copy of the bucket."""
        dl = dlogits.detach().to(torch.float32).contiguous()
        if ent["bwd"] is not None and ent.get("gbuf_ptr") != gbuf.data_ptr():
            ent["bwd"] = None                 # the bucket moved (data parallelism switched on): capture again
        if ent["bwd"] is None:
            ent["gbuf_ptr"] = gbuf.data_ptr()
            ent["dlogits"] = dl.clone()
            torch.cuda.current_stream(dl.device).synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L_.launches()
            with torch.cuda.graph(g, pool=ent["fwd"].pool(), capture_error_mode="thread_local"):
                gbuf.zero_()
                run(ent["part"], ent["dlogits"], gbuf)
                out = gbuf.clone()
            ent.update(bwd=g, gout=out, bwd_launches=L_.launches() - n0)
            L_.add_launches(-ent["bwd_launches"])
        else:
            ent["dlogits"].copy_(dl)
        ent["bwd"].replay()
        L_.add_launches(ent["bwd_launches"])
        return ent["gout"].clone()   # never hand autograd a view of the static buffer (gradient accumulation would alias it)


class SwinAVEEngine:
    def __init__(self, module):
        self.m = module  # the nn.Module that owns the fp32 master parameters (model/Swin_AVE.py)
        self.cfg = module.cfg
        # Micro-batch lanes (STGCMA_LANES=n, default 1): the clips of a step can be split into n groups that run the same
        # kernel sequence on separate CUDA streams (every op is independent per clip), hoping that the compute-bound
        # kernels of one lane (global cross-modal attention, fused MLP: no DRAM traffic) overlap the HBM-bound kernels
        # of another.  Measured on B200 at 32 clips: 393.8 (1 lane) / 392.7 (2) / 387.2 (3) clips/s -- every kernel
        # already fills all 148 SMs, so a second stream only runs in the tails.  Kept as an option, off by default.
        self.lanes = max(1, int(os.environ.get("STGCMA_LANES", "1")))
        self.arenas = [_Arena()]
        self._lane_streams = []
        self.frozen = None          # cache of frozen weights in compute dtype
        self.frozen_key = None
        self.train_pack = None      # packed trainable weights + table
        self.train_key = None
        self.generation = 0
        self.dp = _DpExchange()     # mean of the gradient bucket over the ranks when torch.distributed is initialised
        # dW of the adapters (gemm_tn) is off the dX critical path: it runs on a side stream, overlapping the next
        # kernels of the main chain, and is joined before its operand buffers are rewritten (backward())
        self.side_stream_dw = os.environ.get("STGCMA_SIDE_DW", "1") != "0"
        self._side = {}
        self.graphs = _StepGraphs()

    @property
    def dp_allreduce(self):
        return self.dp.enabled

    @dp_allreduce.setter
    def dp_allreduce(self, v):
        self.dp.enabled = bool(v)

    @property
    def arena(self):
        """All persistent activation buffers (bench.py reports their size)."""
        class _All:
            def __init__(self, arenas):
                self.arenas = arenas

            def bytes(self):
                return sum(a.bytes() for a in self.arenas)

            def clear(self):
                for a in self.arenas:
                    a.clear()
        return _All(self.arenas)

    # ------------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def stream_dtype(adt):
        """dtype of the residual stream x / dx.  fp32 in both modes: the reference under the caller's autocast
        also keeps the residual in fp32 (LayerNorm and the residual adds run in fp32, only Linear / bmm
        outputs are half), and a bf16 stream re-rounds the sum once per sub-block, which alone puts
        adapter gradients 3-5 % off after 12 blocks (measured)."""
        return torch.float32 if adt == torch.bfloat16 else adt

    def _blocks(self):
        for s, layer in enumerate(self.m.layers):
            for i, blk in enumerate(layer.blocks):
                yield s, i, blk

    def _frozen_params(self):
        m = self.m
        ps = [m.patch_embed.proj.weight, m.patch_embed.proj.bias, m.patch_embed.norm.weight, m.patch_embed.norm.bias,
              m.patch_embed_audio.proj.weight, m.patch_embed_audio.proj.bias, m.patch_embed_audio.norm.weight,
              m.patch_embed_audio.norm.bias, m.norm.weight, m.norm.bias]
        for _, _, b in self._blocks():
            ps += [b.norm1.weight, b.norm1.bias, b.norm2.weight, b.norm2.bias, b.attn.qkv.weight, b.attn.qkv.bias,
                   b.attn.proj.weight, b.attn.proj.bias, b.attn.relative_position_bias_table, b.mlp.fc1.weight,
                   b.mlp.fc1.bias, b.mlp.fc2.weight, b.mlp.fc2.bias]
        for layer in self.m.layers:
            if layer.downsample is not None:
                ps += [layer.downsample.norm.weight, layer.downsample.norm.bias, layer.downsample.reduction.weight]
        return ps

    def trainable_params(self):
        """Ordered list of (name, parameter) of everything the engine produces gradients for."""
        out = []
        for s, i, b in self._blocks():
            p = f"layers.{s}.blocks.{i}."
            out += [(p + "gate_v", b.gate_v), (p + "gate_a", b.gate_a)]
            if b.t_attn:
                out += [(p + "attn.temporal_position_bias_table", b.attn.temporal_position_bias_table),
                        (p + "attn.temporal_position_bias_table_audio", b.attn.temporal_position_bias_table_audio)]
            names = (["T_Adapter", "T_Adapter_Audio"] if b.t_attn else []) + \
                ["S_Adapter2", "S_Adapter2_Audio", "S_Adapter", "S_Adapter_Audio"]
            for nm in names:
                ad = getattr(b, nm)
                out += [(p + nm + ".D_fc1.weight", ad.D_fc1.weight), (p + nm + ".D_fc1.bias", ad.D_fc1.bias),
                        (p + nm + ".D_fc2.weight", ad.D_fc2.weight), (p + nm + ".D_fc2.bias", ad.D_fc2.bias)]
        h = self.m.mlp_head
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
        m, fz = self.m, {}

        def w(t):   # GEMM operand in compute dtype
            return t.detach().to(device=dev, dtype=adt).contiguous()

        def wt(t):  # transposed operand for dX = dY . W
            return t.detach().to(device=dev, dtype=adt).t().contiguous()

        def f(t):
            return t.detach().to(device=dev, dtype=torch.float32).contiguous()

        for nm in ("patch_embed", "patch_embed_audio"):
            pe = getattr(m, nm)
            C = pe.proj.weight.shape[0]
            fz[nm] = dict(w=f(pe.proj.weight).reshape(C, -1).contiguous(), b=f(pe.proj.bias), g=f(pe.norm.weight),
                          beta=f(pe.norm.bias))
        for s, i, b in self._blocks():
            fz[(s, i)] = dict(
                n1g=f(b.norm1.weight), n1b=f(b.norm1.bias), n2g=f(b.norm2.weight), n2b=f(b.norm2.bias),
                qkv_w=w(b.attn.qkv.weight), qkv_wt=wt(b.attn.qkv.weight), qkv_b=f(b.attn.qkv.bias),
                proj_w=w(b.attn.proj.weight), proj_wt=wt(b.attn.proj.weight), proj_b=f(b.attn.proj.bias),
                rpb=f(b.attn.relative_position_bias_table),
                fc1_w=w(b.mlp.fc1.weight), fc1_wt=wt(b.mlp.fc1.weight), fc1_b=f(b.mlp.fc1.bias),
                fc2_w=w(b.mlp.fc2.weight), fc2_wt=wt(b.mlp.fc2.weight), fc2_b=f(b.mlp.fc2.bias))
        for s, layer in enumerate(m.layers):
            if layer.downsample is not None:
                d = layer.downsample
                fz[("merge", s)] = dict(g=f(d.norm.weight), b=f(d.norm.bias), w=w(d.reduction.weight),
                                        wt=wt(d.reduction.weight))
        fz["norm"] = dict(g=f(m.norm.weight), b=f(m.norm.bias))
        self.frozen, self.frozen_key = fz, key
        return fz

    def _build_train_pack(self, dev, adt):
        """Compute-dtype copies of the trainable parameters + the device table that refreshes them."""
        tp = self.trainable_params()
        key = (str(dev), adt, tuple(p.data_ptr() for _, p in tp))
        if self.train_pack is not None and self.train_key == key:
            return self.train_pack
        for _, p in tp:
            if p.device != dev or p.dtype != torch.float32:
                raise L_.StgcmaError("trainable parameters must be fp32 tensors on the compute device")
        cfg = self.cfg
        entries, pk = [], {}

        def z(*shape, dtype=adt):
            return torch.zeros(*shape, dtype=dtype, device=dev)

        def add(src, dst, rows, cols, src_ld, dst_ld, transpose=False, f32=False):
            entries.append(L_.pack_entry(src, dst, rows, cols, src_ld, dst_ld, transpose, f32))

        for s, i, b in self._blocks():
            C, r = cfg.stage_dim(s), cfg.stage_r(s)
            rp = _rup(r, _EPS_PAD)
            d = dict(rp=rp, r=r)
            gate = z(2, dtype=torch.float32)
            add(b.gate_v.data, gate[0:1], 1, 1, 1, 1, f32=True)
            add(b.gate_a.data, gate[1:2], 1, 1, 1, 1, f32=True)
            d["gate"] = gate
            if b.t_attn:
                T2, heads = b.attn.temporal_position_bias_table.shape
                tbl = z(2, T2, heads, dtype=torch.float32)
                add(b.attn.temporal_position_bias_table.data, tbl[0], T2, heads, heads, heads, f32=True)
                add(b.attn.temporal_position_bias_table_audio.data, tbl[1], T2, heads, heads, heads, f32=True)
                d["tbl"] = tbl
            groups = ([("T", "T_Adapter", "T_Adapter_Audio")] if b.t_attn else []) + \
                [("S2", "S_Adapter2", "S_Adapter2_Audio"), ("S", "S_Adapter", "S_Adapter_Audio")]
            for tag, nv, na in groups:
                w1, b1 = z(2, rp, C), z(2, rp, dtype=torch.float32)      # fc1:  [rp, C] (rows >= r zero)
                w2, b2 = z(2, C, rp), z(2, C, dtype=torch.float32)       # fc2:  [C, rp] (cols >= r zero)
                w1t, w2t = z(2, C, rp), z(2, rp, C)                      # transposed operands for dX
                for g, nm in enumerate((nv, na)):
                    ad = getattr(b, nm)
                    add(ad.D_fc1.weight.data, w1[g], r, C, C, C)
                    add(ad.D_fc1.weight.data, w1t[g], r, C, C, rp, transpose=True)
                    add(ad.D_fc1.bias.data, b1[g], 1, r, r, rp, f32=True)
                    add(ad.D_fc2.weight.data, w2[g], C, r, r, rp)
                    add(ad.D_fc2.weight.data, w2t[g], C, r, r, C, transpose=True)
                    add(ad.D_fc2.bias.data, b2[g], 1, C, C, C, f32=True)
                d[tag] = dict(w1=w1, b1=b1, w2=w2, b2=b2, w1t=w1t, w2t=w2t)
            pk[(s, i)] = d
        h0, h2 = self.m.mlp_head[0], self.m.mlp_head[2]
        nl = h2.weight.shape[0]
        lp = _rup(nl, 16)
        F2, Hd = h0.weight.shape[1], h0.weight.shape[0]
        # the head runs in fp32 in both modes (tiny: B*T rows): its tensors sit at the root of the backward,
        # where a rounding error is shared by every gradient downstream
        f32 = torch.float32
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
        table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        pk["_table"], pk["_n"] = table, len(entries)
        # flat gradient bucket: offsets in trainable_params() order; (gate_v, gate_a) and the two
        # temporal tables of a block are adjacent so one kernel writes both.
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
        dev = v.device
        adt = m.compute_dtype
        if need_grad:
            self._check_frozen()
        fz = self._build_frozen(dev, adt)
        pk = self._build_train_pack(dev, adt)
        self.generation += 1
        B, Cin, T, H, W = v.shape
        assert Cin == 3 and T == cfg.num_frames and H == cfg.img_size and W == cfg.img_size, "video must be [B,3,T,H,W]"
        assert tuple(a.shape) == (B, T, H, W), "audio must be [B,T,H,W] (Swin path)"
        a32 = a.detach().to(torch.float32).contiguous()
        v32 = v.detach().to(torch.float32).contiguous()
        lanes = min(self.lanes, B) if dev.type == "cuda" else 1
        while len(self.arenas) < lanes:
            self.arenas.append(_Arena())
        if lanes == 1 and self.graphs.wanted(B, dev):
            return self._forward_graphed(a32, v32, mode, need_grad, fz, pk)
        L_.pack_params(pk["_table"], pk["_n"], L_.dt(adt))
        if lanes == 1:
            logits, part = self._forward_part(a32, v32, mode, self.arenas[0], fz, pk)
            return logits, dict(gen=self.generation, mode=mode, dev=dev, parts=[part], rows=[(0, B * T)])
        main = torch.cuda.current_stream(dev)
        while len(self._lane_streams) < lanes:
            self._lane_streams.append(torch.cuda.Stream(device=dev))
        ready = torch.cuda.Event()
        ready.record(main)
        bounds = [(B * k) // lanes for k in range(lanes + 1)]
        outs, parts, rows = [], [], []
        for k in range(lanes):
            b0, b1 = bounds[k], bounds[k + 1]
            st = self._lane_streams[k]
            with torch.cuda.stream(st):
                st.wait_event(ready)
                lg, part = self._forward_part(a32[b0:b1], v32[b0:b1], mode, self.arenas[k], fz, pk)
                lg.record_stream(main)
                done = torch.cuda.Event()
                done.record(st)
            main.wait_event(done)
            outs.append(lg)
            parts.append(part)
            rows.append((b0 * T, b1 * T))
        a32.record_stream(self._lane_streams[0]); v32.record_stream(self._lane_streams[0])
        for st in self._lane_streams[1:lanes]:
            a32.record_stream(st); v32.record_stream(st)
        return torch.cat(outs, 0), dict(gen=self.generation, mode=mode, dev=dev, parts=parts, rows=rows)

    # graph_mode / _graphs: test and script hooks (tests/test_step_gpu.py, scripts/profile_step.py)
    @property
    def graph_mode(self):
        return self.graphs.mode

    @graph_mode.setter
    def graph_mode(self, v):
        self.graphs.mode = v

    @property
    def _graphs(self):
        return self.graphs.entries

    def _forward_graphed(self, a32, v32, mode, need_grad, fz, pk):
        m = self.m
        dev, adt = v32.device, m.compute_dtype
        B, T = v32.shape[0], v32.shape[2]
        key = (mode, bool(need_grad), m.training, m.mlp_head[1].training, tuple(v32.shape), dev.index, adt, self.side_stream_dw)

        def run(a, v):
            L_.pack_params(pk["_table"], pk["_n"], L_.dt(adt))
            return self._forward_part(a, v, mode, self.arenas[0], fz, pk)
        logits, part, ent = self.graphs.forward(key, fz, pk, a32, v32, run)
        return logits, dict(gen=self.generation, mode=mode, dev=dev, rows=[(0, B * T)], parts=[part], graph=ent)

    def saved_masks(self, sv):
        """DropPath masks {(stage, block): flat [2, B, T, L]} and the head-dropout keep mask [B*T, hidden] of a forward,
        merged over the micro-batch lanes (test hook: the oracle replays them)."""
        parts = sv["parts"]
        if len(parts) == 1:
            return dict(dp=parts[0]["dp"], keepmask=parts[0]["keepmask"])
        T = self.cfg.num_frames
        dp = {}
        for key in parts[0]["dp"]:
            if parts[0]["dp"][key] is None:
                dp[key] = None
                continue
            dp[key] = torch.cat([p["dp"][key].view(2, p["B"], T, -1) for p in parts], 1).reshape(-1)
        keep = None if parts[0]["keepmask"] is None else torch.cat([p["keepmask"] for p in parts], 0)
        return dict(dp=dp, keepmask=keep)

    def _forward_part(self, a32, v32, mode, ar, fz, pk):
        """The forward schedule over the clips of one lane (a32 [B,T,H,W], v32 [B,3,T,H,W], fp32, contiguous)."""
        cfg, m = self.cfg, self.m
        dev = v32.device
        adt = m.compute_dtype
        fusion = mode == "fusion"
        B, Cin, T, H, W = v32.shape
        BT = B * T
        frames2 = 2 * BT
        training = m.training
        sdt = self.stream_dtype(adt)   # residual stream x: fp32 even in bf16 mode
        sv = dict(B=B, mode=mode, dp={}, dev=dev, adt=adt, sdt=sdt, arena=ar)

        def buf(name, *shape, dtype=adt, zero=False):
            return ar.get(name, shape, dtype, dev, zero)

        # ---- patch embed (video rows, then audio rows) ----
        C0, R0 = cfg.stage_dim(0), cfg.stage_res(0)
        L0 = R0 * R0
        x = buf("x_in", 2 * BT * L0, C0, dtype=sdt)
        pe, pa = fz["patch_embed"], fz["patch_embed_audio"]
        L_.patch_embed_fwd(v32, pe["w"], pe["b"], pe["g"], pe["beta"], x[:BT * L0], B, 3, T, H, W, C0)
        L_.patch_embed_fwd(a32, pa["w"], pa["b"], pa["g"], pa["beta"], x[BT * L0:], B, 1, T, H, W, C0)

        for s in range(cfg.num_stages):
            C, R, heads = cfg.stage_dim(s), cfg.stage_res(s), cfg.num_heads[s]
            Ls = R * R
            M = frames2 * Ls
            for i in range(cfg.depths[s]):
                blk = m.layers[s].blocks[i]
                f, p = fz[(s, i)], pk[(s, i)]
                rp = p["rp"]
                ws, shift = cfg.block_window_shift(s, i)
                k = f"b{s}_{i}_"
                # ---------------- (T) temporal attention + T_Adapter ----------------
                if blk.t_attn:
                    xn = buf(f"xn{s}", M, C)
                    mean, rstd = buf(k + "t_mean", M, dtype=torch.float32), buf(k + "t_rstd", M, dtype=torch.float32)
                    L_.layernorm_fwd(x, f["n1g"], f["n1b"], xn, mean, rstd, M, C)
                    qkv_t = buf(k + "qkv_t", M, 3 * C)
                    L_.gemm(xn, f["qkv_w"], qkv_t, M, 3 * C, C, bias=f["qkv_b"])
                    o = buf(f"o{s}", M, C)
                    L_.temporal_attn_fwd(qkv_t, p["tbl"], o, 2, B, T, Ls, heads)
                    rt = buf(k + "rt", M, C)
                    L_.gemm(o, f["proj_w"], rt, M, C, C, bias=f["proj_b"])
                    hpre, h = buf(k + "t_hpre", M, rp), buf(k + "t_h", M, rp)
                    A = p["T"]
                    L_.gemm(rt, A["w1"], h, M, rp, C, bias=A["b1"], out_pre=hpre, groups=2, epi=L_.EPI_GELU)
                    dp = None
                    if training and blk.drop_path_rate > 0.0:
                        keep = 1.0 - blk.drop_path_rate
                        mask = torch.empty(2, B, 1, Ls, device=dev, dtype=torch.float32).bernoulli_(keep).div_(keep)
                        dp = mask.expand(2, B, T, Ls).contiguous().view(-1)
                    sv["dp"][(s, i)] = dp
                    x1 = buf(k + "x1", M, C, dtype=sdt)
                    L_.gemm(h, A["w2"], x1, M, C, rp, bias=A["b2"], res1=x, rowscale=dp, groups=2)
                    sv[k + "x0"] = x
                    x = x1
                # ---------------- (S) window attention + windowed cross-modal adapter ----------------
                xn = buf(f"xn{s}", M, C)
                mean, rstd = buf(k + "s_mean", M, dtype=torch.float32), buf(k + "s_rstd", M, dtype=torch.float32)
                L_.layernorm_fwd(x, f["n1g"], f["n1b"], xn, mean, rstd, M, C)
                qkv = buf(k + "qkv", M, 3 * C)
                L_.gemm(xn, f["qkv_w"], qkv, M, 3 * C, C, bias=f["qkv_b"])
                o = buf(f"o{s}", M, C)
                L_.window_attn_fwd(qkv, f["rpb"], o, frames2, R, ws, shift, heads)
                at = buf(k + "attn", M, C)
                L_.gemm(o, f["proj_w"], at, M, C, C, bias=f["proj_b"])
                hpre, h = buf(k + "s_hpre", M, rp), buf(k + "s_h", M, rp)
                A = p["S2"]
                L_.gemm(at, A["w1"], h, M, rp, C, bias=A["b1"], out_pre=hpre, groups=2, epi=L_.EPI_GELU)
                if fusion:
                    h2, att = buf(k + "s_h2", M, rp), buf(k + "s_att", M, rp)
                    L_.xmodal_window_fwd(h, p["gate"], h2, att, BT, R, ws, shift, rp, r=p["r"])
                else:
                    h2 = h
                x2 = buf(k + "x2", M, C, dtype=sdt)
                L_.gemm(h2, A["w2"], x2, M, C, rp, bias=A["b2"], res1=x, res2=at, groups=2)
                sv[k + "x1"] = x
                x = x2
                # ---------------- (G) frozen MLP + global cross-modal adapter ----------------
                mean, rstd = buf(k + "g_mean", M, dtype=torch.float32), buf(k + "g_rstd", M, dtype=torch.float32)
                hid = f["fc1_w"].shape[0]
                y = buf(k + "y", M, C)
                hpre, h = buf(k + "g_hpre", M, rp), buf(k + "g_h", M, rp)
                A = p["S"]
                if L_.mlp_fused_ok(C, hid, rp, adt):
                    # one kernel: LN2 -> fc1 -> GELU -> fc2 -> adapter fc1 (+GELU); the 4C-wide hidden stays on chip
                    L_.mlp_fwd(x, f["n2g"], f["n2b"], f["fc1_w"], f["fc1_b"], f["fc2_w"], f["fc2_b"], A["w1"], A["b1"], y,
                               hpre, h, mean, rstd, M, C, hid, rp)
                else:
                    xn = buf(f"xn{s}", M, C)
                    L_.layernorm_fwd(x, f["n2g"], f["n2b"], xn, mean, rstd, M, C)
                    fpre, fact = buf(k + "fpre", M, hid), buf(f"fact{s}", M, hid)
                    # fpre holds GELU'(fc1(xn)) (derivative-saving epilogue): the backward of the activation is a multiply
                    L_.gemm(xn, f["fc1_w"], fact, M, hid, C, bias=f["fc1_b"], out_pre=fpre, epi=L_.EPI_GELU_G)
                    L_.gemm(fact, f["fc2_w"], y, M, C, hid, bias=f["fc2_b"])
                    L_.gemm(y, A["w1"], h, M, rp, C, bias=A["b1"], out_pre=hpre, groups=2, epi=L_.EPI_GELU)
                if fusion:
                    h2, att = buf(k + "g_h2", M, rp), buf(k + "g_att", M, rp)
                    lse = buf(k + "g_lse", 2 * BT * Ls, dtype=torch.float32)
                    L_.xmodal_global_fwd(h, p["gate"], h2, att, lse, BT, Ls, rp, r=p["r"])
                else:
                    h2 = h
                x3 = buf(k + "x3", M, C, dtype=sdt)
                L_.gemm(h2, A["w2"], x3, M, C, rp, bias=A["b2"], res1=x, res2=y, groups=2)
                sv[k + "x2"] = x
                x = x3
            if s < cfg.num_stages - 1:
                # ---------------- PatchMerging: 2x2 gather + LN(4C) + Linear(4C -> 2C) ----------------
                mg = fz[("merge", s)]
                Mn = M // 4
                xm = buf(f"xm{s}", Mn, 4 * C)
                mean, rstd = buf(f"m{s}_mean", Mn, dtype=torch.float32), buf(f"m{s}_rstd", Mn, dtype=torch.float32)
                L_.layernorm_fwd(x, mg["g"], mg["b"], xm, mean, rstd, Mn, 4 * C, merge=1, R=R)
                xnext = buf(f"m{s}_out", Mn, 2 * C, dtype=sdt)
                L_.gemm(xm, mg["w"], xnext, Mn, 2 * C, 4 * C)
                sv[f"m{s}_in"] = x
                x = xnext
        # ---------------- head ----------------
        s = cfg.num_stages - 1
        C, R = cfg.stage_dim(s), cfg.stage_res(s)
        Ls = R * R
        M = frames2 * Ls
        hd = pk["head"]
        f32 = torch.float32
        xn = buf("h_xn", M, C, dtype=f32)
        mean, rstd = buf("h_mean", M, dtype=torch.float32), buf("h_rstd", M, dtype=torch.float32)
        L_.layernorm_fwd(x, fz["norm"]["g"], fz["norm"]["b"], xn, mean, rstd, M, C)
        feat = buf("h_feat", BT, 2 * C, dtype=f32)
        L_.head_pool_fwd(xn, feat, BT, Ls, C)
        Hd = hd["w0"].shape[0]
        hh = buf("h_hidden", BT, Hd, dtype=f32)
        drop = m.mlp_head[1]
        keepmask = None
        if drop.training and drop.p > 0.0:
            keepmask = torch.empty(BT, Hd, device=dev, dtype=torch.float32).bernoulli_(1.0 - drop.p).div_(1.0 - drop.p)
            L_.gemm(feat, hd["w0"], hh, BT, Hd, 2 * C, bias=hd["b0"], aux=keepmask, epi=L_.EPI_MULAUX)
        else:
            L_.gemm(feat, hd["w0"], hh, BT, Hd, 2 * C, bias=hd["b0"])
        sv["keepmask"] = keepmask
        lp, nl = hd["lp"], hd["nl"]
        logits_p = buf("h_logits", BT, lp, dtype=f32)
        L_.gemm(hh, hd["w2"], logits_p, BT, lp, Hd, bias=hd["b2"])
        sv["x_final"] = x
        logits = logits_p[:, :nl].clone()
        return logits, sv

    # ------------------------------------------------------------------------------------------
    # backward
    # ------------------------------------------------------------------------------------------
    def backward(self, sv, dlogits):
        if sv["gen"] != self.generation:
            raise L_.StgcmaError("backward() of a stale forward: the activation arena holds one step at a time "
                                 "(run forward -> backward -> forward ...)")
        dev = sv["dev"]
        pk = self.train_pack
        parts = sv["parts"]
        gbuf = self.dp.bucket(pk["_gsize"], dev)
        if gbuf is None:
            gbuf = self.arenas[0].get("grad_bucket", (pk["_gsize"],), torch.float32, dev)
        out = None
        if sv.get("graph") is not None:
            out = self.graphs.backward(sv["graph"], dlogits, gbuf, self._backward_part)
        else:
            gbuf.zero_()
        if out is not None:
            pass
        elif len(parts) == 1:
            self._backward_part(parts[0], dlogits, gbuf)
        else:
            main = torch.cuda.current_stream(dev)
            ready = torch.cuda.Event()
            ready.record(main)
            for k, (part, (r0, r1)) in enumerate(zip(parts, sv["rows"])):
                st = self._lane_streams[k]
                with torch.cuda.stream(st):
                    st.wait_event(ready)
                    self._backward_part(part, dlogits[r0:r1], gbuf)
                    done = torch.cuda.Event()
                    done.record(st)
                main.wait_event(done)
                dlogits.record_stream(st)
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
        """Backward schedule over one lane's clips; weight gradients are ACCUMULATED (atomics) into `gbuf`."""
        cfg, m = self.cfg, self.m
        ar = sv["arena"]
        dev, adt, sdt, B = sv["dev"], sv["adt"], sv["sdt"], sv["B"]
        fz, pk = self.frozen, self.train_pack
        fusion = sv["mode"] == "fusion"
        T = cfg.num_frames
        BT = B * T
        frames2 = 2 * BT

        def buf(name, *shape, dtype=adt, zero=False):
            return ar.get(name, shape, dtype, dev, zero)

        def slot(name):
            return ar.slots[name]

        goffs = pk["_goffs"]

        def g(name):
            off, shape = goffs[name]
            return gbuf[off:off + math.prod(shape)]

        # ---------------- head ----------------
        s = cfg.num_stages - 1
        C, R = cfg.stage_dim(s), cfg.stage_res(s)
        Ls = R * R
        M = frames2 * Ls
        hd = pk["head"]
        lp, nl = hd["lp"], hd["nl"]
        Hd = hd["w0"].shape[0]
        f32 = torch.float32
        dlog = buf("d_logits", BT, lp, dtype=f32, zero=True)
        dlog[:, :nl].copy_(dlogits)
        hh, feat = slot("h_hidden"), slot("h_feat")
        L_.gemm_tn(dlog, hh, g("mlp_head.2.weight"), BT, nl, Hd, lda=lp, ldb=Hd, ldo=Hd, csumA=g("mlp_head.2.bias"))
        dh = buf("d_hidden", BT, Hd, dtype=f32)
        if sv["keepmask"] is not None:
            L_.gemm(dlog, hd["w2t"], dh, BT, Hd, lp, aux=sv["keepmask"], epi=L_.EPI_MULAUX)
        else:
            L_.gemm(dlog, hd["w2t"], dh, BT, Hd, lp)
        L_.gemm_tn(dh, feat, g("mlp_head.0.weight"), BT, Hd, 2 * C, lda=Hd, ldb=2 * C, ldo=2 * C, csumA=g("mlp_head.0.bias"))
        dfeat = buf("d_feat", BT, 2 * C, dtype=f32)
        L_.gemm(dh, hd["w0t"], dfeat, BT, 2 * C, Hd)
        dxn = buf("d_hxn", M, C, dtype=f32)
        L_.head_pool_bwd(dfeat, dxn, BT, Ls, C)
        # dx = gradient of the residual stream (stream dtype); dxb = its activation-dtype copy, written by the
        # same kernel, which is what the GEMMs consume as an operand
        dx, dxb = buf(f"dB{s}", M, C, dtype=sdt), buf(f"dBb{s}", M, C)
        L_.layernorm_bwd(dxn, sv["x_final"], slot("h_mean"), slot("h_rstd"), fz["norm"]["g"], None, dx, M, C, dx_copy=dxb)

        side = None
        main = torch.cuda.current_stream(dev) if dev.type == "cuda" else None   # (the CPU schedule tests mock the kernels)
        if self.side_stream_dw and main is not None:
            key = main.cuda_stream
            if key not in self._side:
                self._side[key] = torch.cuda.Stream(device=dev)
            side = self._side[key]
        pending = []   # events of side-stream dW work whose operand buffers the main chain has not rewritten yet

        def join_side():
            while pending:
                main.wait_event(pending.pop())

        def adapter_grads(prefix_names, d_out, h_in, d_hpre, x_in, M, C, r, rp, rowscale):
            """dW2/db2 from (d_out, h_in); dW1/db1 from (d_hpre, x_in), per modality."""
            Mg = M // 2

            def run():
                for gi, nm in enumerate(prefix_names):
                    rs = None if rowscale is None else rowscale[gi * Mg:(gi + 1) * Mg]
                    L_.gemm_tn(d_out[gi * Mg:], h_in[gi * Mg:], g(nm + ".D_fc2.weight"), Mg, C, r, lda=C, ldb=rp, ldo=r,
                               rowscale=rs, csumA=g(nm + ".D_fc2.bias"))
                    L_.gemm_tn(d_hpre[gi * Mg:], x_in[gi * Mg:], g(nm + ".D_fc1.weight"), Mg, r, C, lda=rp, ldb=C, ldo=C,
                               csumA=g(nm + ".D_fc1.bias"))
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

        for s in reversed(range(cfg.num_stages)):
            C, R, heads = cfg.stage_dim(s), cfg.stage_res(s), cfg.num_heads[s]
            Ls = R * R
            M = frames2 * Ls
            if s < cfg.num_stages - 1:
                # dx currently is the gradient of the merged output [M/4, 2C]
                mg = fz[("merge", s)]
                Mn = M // 4
                dxm = buf(f"dxm{s}", Mn, 4 * C)
                L_.gemm(dxb, mg["wt"], dxm, Mn, 4 * C, 2 * C)
                dx, dxb = buf(f"dB{s}", M, C, dtype=sdt), buf(f"dBb{s}", M, C)
                L_.layernorm_bwd(dxm, sv[f"m{s}_in"], slot(f"m{s}_mean"), slot(f"m{s}_rstd"), mg["g"], None, dx, Mn, 4 * C,
                                 merge=1, R=R, dx_copy=dxb)
            for i in reversed(range(cfg.depths[s])):
                blk = m.layers[s].blocks[i]
                f, p = fz[(s, i)], pk[(s, i)]
                r, rp = p["r"], p["rp"]
                ws, shift = cfg.block_window_shift(s, i)
                k = f"b{s}_{i}_"
                pfx = f"layers.{s}.blocks.{i}."
                first_block = (s == 0 and i == 0)
                hid = f["fc1_w"].shape[0]
                # [gate_v, gate_a] are adjacent in the bucket: one kernel accumulates both
                assert goffs[pfx + "gate_a"][0] == goffs[pfx + "gate_v"][0] + 1
                dgate2 = gbuf[goffs[pfx + "gate_v"][0]:goffs[pfx + "gate_v"][0] + 2]
                # ---------------- (G) backward: dx = dX3 ----------------
                join_side()   # the previous sub-block's dW kernels read dhpre / dxb, rewritten below
                A = p["S"]
                y = slot(k + "y")
                dh2 = buf(f"dh{s}", M, rp)
                dhpre = buf(f"dhpre{s}", M, rp)
                if fusion:
                    L_.gemm(dxb, A["w2t"], dh2, M, rp, C, groups=2)
                    work = buf(f"xwork{s}", 2 * BT * Ls, dtype=torch.float32)
                    L_.xmodal_global_bwd(slot(k + "g_h"), slot(k + "g_hpre"), p["gate"], slot(k + "g_att"), slot(k + "g_lse"),
                                         dh2, dhpre, dgate2, work, BT, Ls, rp, r=r)
                    h2 = slot(k + "g_h2")
                else:
                    L_.gemm(dxb, A["w2t"], dhpre, M, rp, C, groups=2, aux=slot(k + "g_hpre"), epi=L_.EPI_DGELU)
                    h2 = slot(k + "g_h")
                adapter_grads((pfx + "S_Adapter", pfx + "S_Adapter_Audio"), dxb, h2, dhpre, y, M, C, r, rp, None)
                dxn = buf(f"dC{s}", M, C)
                if L_.mlp_fused_ok(C, hid, rp, adt):
                    # one kernel: dy = dhpre W1a + dxb -> (dy W2) * gelu'(recomputed fc1(LN2(x))) -> . W1
                    L_.mlp_bwd(dhpre, dxb, sv[k + "x2"], f["n2g"], f["n2b"], slot(k + "g_mean"), slot(k + "g_rstd"), A["w1t"],
                               f["fc1_w"], f["fc1_b"], f["fc2_wt"], f["fc1_wt"], dxn, M, C, hid, rp)
                else:
                    dy = buf(f"dA{s}", M, C)
                    L_.gemm(dhpre, A["w1t"], dy, M, C, rp, res1=dxb, groups=2)
                    dfpre = buf(f"dfpre{s}", M, hid)
                    L_.gemm(dy, f["fc2_wt"], dfpre, M, hid, C, aux=slot(k + "fpre"), epi=L_.EPI_MULAUX)
                    L_.gemm(dfpre, f["fc1_wt"], dxn, M, C, hid)
                # the three sub-blocks rotate over three buffer pairs (G -> dE, S -> dB, T -> dD): the side-stream dW kernels
                # launched above still read dxb, which after a (T) sub-block IS dDb -- writing this sub-block's result there
                # would race with them (they are joined only at the start of the next sub-block)
                dx2, dx2b = buf(f"dE{s}", M, C, dtype=sdt), buf(f"dEb{s}", M, C)
                L_.layernorm_bwd(dxn, sv[k + "x2"], slot(k + "g_mean"), slot(k + "g_rstd"), f["n2g"], dx, dx2, M, C, dx_copy=dx2b)
                # ---------------- (S) backward: dx2 ----------------
                join_side()
                A = p["S2"]
                at = slot(k + "attn")
                if fusion:
                    L_.gemm(dx2b, A["w2t"], dh2, M, rp, C, groups=2)
                    L_.xmodal_window_bwd(slot(k + "s_h"), slot(k + "s_hpre"), p["gate"], slot(k + "s_att"), dh2, dhpre, dgate2,
                                         BT, R, ws, shift, rp, r=r)
                    h2 = slot(k + "s_h2")
                else:
                    L_.gemm(dx2b, A["w2t"], dhpre, M, rp, C, groups=2, aux=slot(k + "s_hpre"), epi=L_.EPI_DGELU)
                    h2 = slot(k + "s_h")
                adapter_grads((pfx + "S_Adapter2", pfx + "S_Adapter2_Audio"), dx2b, h2, dhpre, at, M, C, r, rp, None)
                if first_block and not blk.t_attn:
                    break  # nothing upstream needs a gradient
                dat = buf(f"dA{s}", M, C)
                L_.gemm(dhpre, A["w1t"], dat, M, C, rp, res1=dx2b, groups=2)
                do = buf(f"dC{s}", M, C)
                L_.gemm(dat, f["proj_wt"], do, M, C, C)
                dqkv = buf(f"dqkv{s}", M, 3 * C)
                L_.window_attn_bwd(slot(k + "qkv"), f["rpb"], do, dqkv, frames2, R, ws, shift, heads)
                dxn = buf(f"dA{s}", M, C)
                L_.gemm(dqkv, f["qkv_wt"], dxn, M, C, 3 * C)
                dx1, dx1b = buf(f"dB{s}", M, C, dtype=sdt), buf(f"dBb{s}", M, C)
                L_.layernorm_bwd(dxn, sv[k + "x1"], slot(k + "s_mean"), slot(k + "s_rstd"), f["n1g"], dx2, dx1, M, C, dx_copy=dx1b)
                # ---------------- (T) backward: dx1 ----------------
                if blk.t_attn:
                    join_side()
                    A = p["T"]
                    dp = sv["dp"][(s, i)]
                    L_.gemm(dx1b, A["w2t"], dhpre, M, rp, C, groups=2, rowscale=dp, aux=slot(k + "t_hpre"), epi=L_.EPI_DGELU)
                    adapter_grads((pfx + "T_Adapter", pfx + "T_Adapter_Audio"), dx1b, slot(k + "t_h"), dhpre, slot(k + "rt"),
                                  M, C, r, rp, dp)
                    drt = buf(f"dA{s}", M, C)
                    L_.gemm(dhpre, A["w1t"], drt, M, C, rp, groups=2)
                    do = buf(f"dC{s}", M, C)
                    L_.gemm(drt, f["proj_wt"], do, M, C, C)
                    dqkv = buf(f"dqkv{s}", M, 3 * C)
                    tname = pfx + "attn.temporal_position_bias_table"
                    assert goffs[tname + "_audio"][0] == goffs[tname][0] + math.prod(goffs[tname][1])
                    dtbl = gbuf[goffs[tname][0]:goffs[tname][0] + 2 * math.prod(goffs[tname][1])]
                    L_.temporal_attn_bwd(slot(k + "qkv_t"), p["tbl"], do, dqkv, dtbl, 2, B, T, Ls, heads)
                    if first_block:
                        break
                    dxn = buf(f"dA{s}", M, C)
                    L_.gemm(dqkv, f["qkv_wt"], dxn, M, C, 3 * C)
                    dx0, dx0b = buf(f"dD{s}", M, C, dtype=sdt), buf(f"dDb{s}", M, C)
                    L_.layernorm_bwd(dxn, sv[k + "x0"], slot(k + "t_mean"), slot(k + "t_rstd"), f["n1g"], dx1, dx0, M, C, dx_copy=dx0b)
                    dx, dxb = dx0, dx0b
                else:
                    dx, dxb = dx1, dx1b

        join_side()
