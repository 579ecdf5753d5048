"""`SLatFlowModel`: the structured-latent flow denoiser of the TRELLIS stage in front of the GVF path (SURVEY.md row f1;
reference trellis/models/structured_latent_flow.py:69-262, called by trellis/pipelines/trellis_image_to_3d.py:223-256
through FlowEulerGuidanceIntervalSampler).  Same constructor arguments, state-dict keys and `forward(x, t, cond)` as the
reference class; inference only, CUDA only (a CPU tensor raises).

    input_layer -> [SparseResBlock3d ... , SparseResBlock3d(downsample)] per io stage -> + APE(coords)
    -> num_blocks x ModulatedSparseTransformerCrossBlock (full sparse self-attention, cross-attention to the dense image
       tokens, MLP; adaLN from the timestep) -> [SparseResBlock3d(upsample), SparseResBlock3d ...] on [h | skip]
    -> LayerNorm -> out_layer

Execution (everything below is this library's kernels; torch supplies memory and the integer bookkeeping of the
resampling plan, once per coordinate set):
  * timestep MLP + EVERY modulation vector of the forward (each ResBlock's emb_layers, each block's adaLN_modulation) as
    one GEMV over the concatenated weights (gvf_dit_modulation), like the DiT engine;
  * ResBlock: LayerNorm (+ affine / + modulate) and SiLU fused into one pass that writes the convolution's fp16 operand
    (gvf_ln_mod_act_f16); both 3x3x3 submanifold convolutions are ONE tcgen05 GEMM each whose TMA producer gathers the
    neighbour rows (tile::gather4, gvf_sparse_conv_gemm_f16); conv2's epilogue adds the skip path in place; the
    downsample is gvf_sparse_pool_mean_f16; the upsample block normalises / projects its skip path at the COARSE level
    (row-wise operators commute with the nearest-neighbour gather) and gathers after;
  * transformer blocks on an fp32 residual stream (the reference keeps it in fp16; fp32 is at least as accurate): LN +
    modulate -> fp16 operand, tcgen05 GEMMs with bias / GELU / gate x + residual epilogues, per-head q / k RMS-norm in
    place, dense tcgen05 flash attention per batch entry for the sparse self-attention (an entry's voxels are contiguous
    rows) and for the cross-attention, whose K / V projections of `cond` are computed once per conditioning tensor and
    reused by every sampler step.
Numerics: the reference's `use_fp16` torso (fp16 weights and activations, LayerNorm32 / softmax statistics in fp32); this
engine rounds at the GEMM operands only.  Parity: tests/test_slat_flow_gpu.py against the CPU restatement that is pinned
to the reference's own class (tests/golden/slat_flow_tiny.pt)."""
import math

import torch

from ... import ops
from ...sparse.basic import SparseTensor
from ...sparse.conv import SparseConv3d
from ...sparse.spatial import SparseDownsample, downsample_plan

F16, F32 = torch.float16, torch.float32


def _h(t, dev):
    return t.detach().to(dev, F16).contiguous()


def _f(t, dev):
    return t.detach().to(dev, F32).contiguous()


class SparseResBlock3d:
    """structured_latent_flow.py:16-66.  `mod_off`: column of this block's (scale | shift) pair in the forward's
    modulation table."""

    def __init__(self, channels, emb_channels, out_channels=None, downsample=False, upsample=False, device="cuda"):
        assert not (downsample and upsample), "Cannot downsample and upsample at the same time"
        self.channels, self.emb_channels, self.out_channels = channels, emb_channels, out_channels or channels
        self.downsample, self.upsample = downsample, upsample
        self.conv1 = SparseConv3d(channels, self.out_channels, 3, device=device)
        self.conv2 = SparseConv3d(self.out_channels, self.out_channels, 3, device=device)
        self.device = torch.device(device)
        self.mod_off = 0
        self.conv_mode = "auto"                # "fused" (TMA gather in the GEMM) | "im2col" | "auto" (measured rule below)

    def load_state_dict(self, sd, prefix):
        dev = self.device
        self.n1w, self.n1b = _f(sd[prefix + "norm1.weight"], dev), _f(sd[prefix + "norm1.bias"], dev)
        self.conv1.load_state_dict(sd, prefix + "conv1.")
        self.conv2.load_state_dict(sd, prefix + "conv2.")
        self.emb_w, self.emb_b = sd[prefix + "emb_layers.1.weight"], sd[prefix + "emb_layers.1.bias"]
        if self.upsample:
            # conv1 acts on a nearest-neighbour upsampled tensor: per-tap weights [(k, out), in] for the coarse-level GEMM
            w = sd[prefix + "conv1.conv.weight"].detach()
            self.tap_w = _h(w.reshape(w.shape[0], 27, w.shape[-1]).permute(1, 0, 2).reshape(27 * w.shape[0], w.shape[-1]), dev)
        self.skip_w = self.skip_b = None
        if prefix + "skip_connection.weight" in sd:
            self.skip_w, self.skip_b = _h(sd[prefix + "skip_connection.weight"], dev), _f(sd[prefix + "skip_connection.bias"], dev)
        return self

    def _conv(self, conv, st, a, out=None, residual=False):
        """One submanifold convolution.  Two executions of the same arithmetic (bit-identical): the GEMM whose TMA producer
        gathers the neighbour rows (no im2col operand), or an explicit fp16 im2col operand + the plain GEMM.  On the B200
        the gather form is bound by the issue rate of its `tile::gather4` loads (one 4 x 128 B box per instruction, ~37 ns
        each: 1.2 us per 128 x 64 operand block, tools/slat_flow_bench.py --conv-ab), so `auto` takes im2col while its
        operand stays below 1 GB and keeps the gather form for larger ones."""
        nbr = conv.neighbor_map(st)
        fused_ok = conv.in_channels % 64 == 0
        mode = self.conv_mode
        if mode == "auto":
            mode = "fused" if fused_ok and a.shape[0] * 27 * conv.in_channels * 2 > (1 << 30) else "im2col"
        if mode == "fused" and fused_ok:
            return ops.sparse_conv_gemm(a, nbr, conv.weight, conv.bias, out=out, residual=residual)
        cols = ops.sparse_im2col(a, nbr)
        return ops.gemm(cols, conv.weight, conv.bias, ops.EPI_RESID_F16 if residual else ops.EPI_F16, out=out)

    def forward(self, st: SparseTensor, feats, mod, idx=None):
        """st: the SparseTensor of the OUTPUT level (coordinates, layout, neighbour-map cache); feats fp16 [rows, channels]
        at that level -- or, for the upsample block, at the coarse level with `idx` = the cell of every output row."""
        Co, R = self.out_channels, mod.shape[1]
        a = ops.ln_mod_act(feats, w=self.n1w, b=self.n1b, act=1)                       # silu(norm1(x)), :57-58
        skip = feats if self.skip_w is None else ops.gemm(feats, self.skip_w, self.skip_b, ops.EPI_F16)
        if idx is not None:                                                            # _updown (:56) after the row-wise work
            # conv1 of the upsampled tensor = per-tap products of the COARSE rows (one GEMM) + a gather-sum over the taps
            P = ops.gemm(a, self.tap_w, None, ops.EPI_F32)
            h = ops.sparse_tap_gather_sum(P, self.conv1.neighbor_map(st), idx, self.conv1.bias)
            skip = ops.gather_concat(a=skip, idx=idx)
        else:
            if self.skip_w is None:
                skip = skip.clone()                                                    # conv2 accumulates in place
            h = self._conv(self.conv1, st, a)
        a2 = torch.empty_like(h)
        for b, s in enumerate(st.layout):                                              # norm2 * (1 + scale) + shift, silu (:60-61)
            if s.stop > s.start:
                ops.ln_mod_act(h[s], out=a2[s], scale=mod[b, self.mod_off:self.mod_off + Co],
                               shift=mod[b, self.mod_off + Co:self.mod_off + 2 * Co], mod_stride=R, act=1)
        return self._conv(self.conv2, st, a2, out=skip, residual=True)                 # conv2 + skip_connection(x), :62-63


class _CrossBlock:
    """ModulatedSparseTransformerCrossBlock (trellis/modules/sparse/transformer/modulated.py:83-166), attn_mode full."""

    def __init__(self, sd, p, C, heads, qk_rms, qk_rms_cross, dev):
        self.C, self.H, self.d = C, heads, C // heads
        self.w_qkv, self.b_qkv = _h(sd[p + "self_attn.to_qkv.weight"], dev), _f(sd[p + "self_attn.to_qkv.bias"], dev)
        self.gq = self.gk = None
        if qk_rms_cross:
            raise NotImplementedError("qk_rms_norm_cross is not used by the shipped structured-latent flow checkpoints")
        if qk_rms:
            self.gq = _f(sd[p + "self_attn.q_rms_norm.gamma"], dev).reshape(-1)
            self.gk = _f(sd[p + "self_attn.k_rms_norm.gamma"], dev).reshape(-1)
        self.w_so, self.b_so = _h(sd[p + "self_attn.to_out.weight"], dev), _f(sd[p + "self_attn.to_out.bias"], dev)
        self.n2w, self.n2b = _f(sd[p + "norm2.weight"], dev), _f(sd[p + "norm2.bias"], dev)
        self.w_q, self.b_q = _h(sd[p + "cross_attn.to_q.weight"], dev), _f(sd[p + "cross_attn.to_q.bias"], dev)
        self.w_kv, self.b_kv = _h(sd[p + "cross_attn.to_kv.weight"], dev), _f(sd[p + "cross_attn.to_kv.bias"], dev)
        self.w_co, self.b_co = _h(sd[p + "cross_attn.to_out.weight"], dev), _f(sd[p + "cross_attn.to_out.bias"], dev)
        self.w1, self.b1 = _h(sd[p + "mlp.mlp.0.weight"], dev), _f(sd[p + "mlp.mlp.0.bias"], dev)
        self.w2, self.b2 = _h(sd[p + "mlp.mlp.2.weight"], dev), _f(sd[p + "mlp.mlp.2.bias"], dev)
        self.mod_w, self.mod_b = sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"]
        self.mod_off = 0

    def context_kv(self, cond16):
        """cond16 fp16 [B, L, Cc] -> K / V [B, L, 2, H, d]."""
        B, L, Cc = cond16.shape
        kv = ops.gemm(cond16.view(B * L, Cc), self.w_kv, self.b_kv, ops.EPI_F16)
        return kv.view(B, L, 2, self.H, self.d)

    def forward(self, X, layout, mod, kv, ws):
        C, H, d, R, o = self.C, self.H, self.d, mod.shape[1], self.mod_off
        A, QKV, AO, H1 = ws["A"], ws["QKV"], ws["AO"], ws["H1"]
        scale = 1.0 / math.sqrt(d)
        m = lambda b, i: mod[b, o + i * C:o + (i + 1) * C]                             # shift/scale/gate msa, shift/scale/gate mlp
        spans = [(b, s) for b, s in enumerate(layout) if s.stop > s.start]
        for b, s in spans:
            ops.ln_mod(X[s], out=A[s], shift=m(b, 0), scale=m(b, 1), mod_stride=R)
        ops.gemm(A, self.w_qkv, self.b_qkv, ops.EPI_F16, out=QKV)
        if self.gq is not None:
            ops.rmsnorm_heads_(QKV, H, d, C, self.gq, self.gk)
        q4 = QKV.view(-1, 3, H, d)
        ao = AO.view(-1, H, d)
        for b, s in spans:                                                              # full attention inside a batch entry
            ops.attention(q4[s, 0][None], q4[s, 1][None], q4[s, 2][None], scale, out=ao[s][None])
        for b, s in spans:
            ops.gemm(AO[s], self.w_so, self.b_so, ops.EPI_RESID_F32, out=X[s], gate=m(b, 2), gate_stride=R,
                     rows_per_batch=s.stop - s.start)
        ops.ln_mod(X, out=A, w=self.n2w, b=self.n2b)
        Q = ops.gemm(A, self.w_q, self.b_q, ops.EPI_F16, out=QKV[:, :C])
        qh = Q.view(-1, H, d)
        for b, s in spans:
            ops.attention(qh[s][None], kv[b:b + 1, :, 0], kv[b:b + 1, :, 1], scale, out=ao[s][None])
        ops.gemm(AO, self.w_co, self.b_co, ops.EPI_RESID_F32, out=X)
        for b, s in spans:
            ops.ln_mod(X[s], out=A[s], shift=m(b, 3), scale=m(b, 4), mod_stride=R)
        ops.gemm(A, self.w1, self.b1, ops.EPI_GELU_F16, out=H1)
        for b, s in spans:
            ops.gemm(H1[s], self.w2, self.b2, ops.EPI_RESID_F32, out=X[s], gate=m(b, 5), gate_stride=R,
                     rows_per_batch=s.stop - s.start)


class SLatFlowModel:
    def __init__(self, resolution, in_channels, model_channels, cond_channels, out_channels, num_blocks, num_heads=None,
                 num_head_channels=64, mlp_ratio=4, patch_size=2, num_io_res_blocks=2, io_block_channels=None, pe_mode="ape",
                 use_fp16=False, use_checkpoint=False, use_skip_connection=True, share_mod=False, qk_rms_norm=False,
                 qk_rms_norm_cross=False, device="cuda"):
        if pe_mode != "ape" or share_mod:
            raise NotImplementedError("rope / share_mod are not used by the shipped structured-latent flow checkpoints")
        assert int(math.log2(patch_size)) == math.log2(patch_size), "Patch size must be a power of 2"
        assert math.log2(patch_size) == len(io_block_channels), "Number of IO ResBlocks must match the number of stages"
        self.resolution, self.in_channels, self.model_channels = resolution, in_channels, model_channels
        self.cond_channels, self.out_channels, self.num_blocks = cond_channels, out_channels, num_blocks
        self.num_heads = num_heads or model_channels // num_head_channels
        self.mlp_ratio, self.patch_size, self.num_io_res_blocks = mlp_ratio, patch_size, num_io_res_blocks
        self.io_block_channels, self.use_skip_connection = list(io_block_channels), use_skip_connection
        self.qk_rms_norm, self.qk_rms_norm_cross = qk_rms_norm, qk_rms_norm_cross
        self.dtype = F16                      # the engine always runs the fp16 torso (use_fp16 of the shipped checkpoints)
        self.device = torch.device(device)
        if in_channels > 32 or model_channels > 1024 or model_channels % 64 or (model_channels // self.num_heads) not in (32, 64):
            raise ValueError("unsupported widths (in_channels <= 32, model_channels <= 1024, head dim 32 / 64)")
        C, io, n = model_channels, self.io_block_channels, num_io_res_blocks
        mk = lambda *a, **k: SparseResBlock3d(*a, device=device, **k)
        self.input_blocks = []                 # (block, is_downsample): structured_latent_flow.py:128-146
        for chs, nxt in zip(io, io[1:] + [C]):
            self.input_blocks += [mk(chs, C, out_channels=chs) for _ in range(n - 1)]
            self.input_blocks.append(mk(chs, C, out_channels=nxt, downsample=True))
        k = 2 if use_skip_connection else 1
        self.out_blocks = []                   # :165-183
        for chs, prev in zip(reversed(io), [C] + list(reversed(io[1:]))):
            self.out_blocks.append(mk(prev * k, C, out_channels=chs, upsample=True))
            self.out_blocks += [mk(chs * k, C, out_channels=chs) for _ in range(n - 1)]
        self.blocks = []
        self._kv_cache = {}
        self._ws = {}
        self._loaded = False
        self.use_graphs = False               # True: __call__ replays a captured graph per (coordinates, conditioning)

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd, strict=True):
        dev, C = self.device, self.model_channels
        self.t_w0, self.t_b0 = _h(sd["t_embedder.mlp.0.weight"], dev), _f(sd["t_embedder.mlp.0.bias"], dev)
        self.t_w2, self.t_b2 = _h(sd["t_embedder.mlp.2.weight"], dev), _f(sd["t_embedder.mlp.2.bias"], dev)
        self.in_w, self.in_b = _h(sd["input_layer.weight"], dev), _f(sd["input_layer.bias"], dev)
        for i, blk in enumerate(self.input_blocks):
            blk.load_state_dict(sd, f"input_blocks.{i}.")
        for i, blk in enumerate(self.out_blocks):
            blk.load_state_dict(sd, f"out_blocks.{i}.")
        self.blocks = [_CrossBlock(sd, f"blocks.{i}.", C, self.num_heads, self.qk_rms_norm, self.qk_rms_norm_cross, dev)
                       for i in range(self.num_blocks)]
        # every modulation vector of one forward out of ONE GEMV: rows = [ResBlock (scale | shift) ... | block 6 C ...]
        ws, bs, off = [], [], 0
        for blk in self.input_blocks + self.out_blocks:
            blk.mod_off = off
            ws.append(blk.emb_w)
            bs.append(blk.emb_b)
            off += 2 * blk.out_channels
        for blk in self.blocks:
            blk.mod_off = off
            ws.append(blk.mod_w)
            bs.append(blk.mod_b)
            off += 6 * C
        self.mod_w, self.mod_b, self.R = _h(torch.cat(ws, 0), dev), _f(torch.cat(bs, 0), dev), off
        wo, bo = sd["out_layer.weight"].detach().float().to(dev), sd["out_layer.bias"].detach().float().to(dev)
        pad = (-self.out_channels) % 8
        self.out_w = _h(torch.cat([wo, torch.zeros(pad, wo.shape[1], device=dev)], 0), dev)
        self.out_b = _f(torch.cat([bo, torch.zeros(pad, device=dev)], 0), dev)
        self._kv_cache.clear()
        self._loaded = True
        return self

    # ------------------------------------------------------------------ caches
    def _context_entry(self, cond):
        """Per conditioning tensor: the K / V of every block, computed once (the sampler calls the model `steps` times with
        the same `cond` / `neg_cond`), and the CUDA graphs captured against them.  The entry keeps `cond` alive so that its
        address cannot be recycled; evicting an entry drops its graphs with it."""
        key = (cond.data_ptr(), cond._version, tuple(cond.shape), cond.dtype)
        hit = self._kv_cache.get(key)
        if hit is None:
            if len(self._kv_cache) >= 4:
                self._kv_cache.pop(next(iter(self._kv_cache)))
            c16 = cond.detach().to(self.device, F16).contiguous()
            hit = {"cond": cond, "kv": [blk.context_kv(c16) for blk in self.blocks], "graphs": {}}
            self._kv_cache[key] = hit
        return hit

    def _context(self, cond):
        return self._context_entry(cond)["kv"]

    def reset_conditioning(self):
        self._kv_cache.clear()

    _MAX_WORKSPACES = 4

    def _workspace(self, n):
        """One set of activation buffers per token count (every object has its own).  Captured graphs point into the
        workspace they were recorded with, so the sets are kept in a small LRU and evicting one drops the graphs that were
        captured at its token count -- memory stays bounded over a stream of objects, no graph ever sees freed buffers."""
        ws = self._ws.pop(n, None)
        if ws is None:
            while len(self._ws) >= self._MAX_WORKSPACES:
                gone = next(iter(self._ws))
                del self._ws[gone]
                for ent in self._kv_cache.values():
                    for k in [k for k, g in ent["graphs"].items() if g[5] == gone]:
                        del ent["graphs"][k]
            C, dev = self.model_channels, self.device
            ws = dict(A=torch.empty((n, C), dtype=F16, device=dev), QKV=torch.empty((n, 3 * C), dtype=F16, device=dev),
                      AO=torch.empty((n, C), dtype=F16, device=dev),
                      H1=torch.empty((n, int(C * self.mlp_ratio)), dtype=F16, device=dev))
        self._ws[n] = ws                                  # most recently used last
        return ws

    # ------------------------------------------------------------------ CUDA-graph replay of one call
    @torch.no_grad()
    def forward_graphed(self, x: SparseTensor, t: torch.Tensor, cond: torch.Tensor) -> SparseTensor:
        """forward() replayed from a CUDA graph (one call = ~450 launches; the sampler repeats it 25 x 2 times on the same
        coordinates and conditioning).  Keyed on the coordinate tensor and the conditioning entry; inputs are copied into
        the graph's static buffers, the result is copied out."""
        if not (x.feats.is_cuda and cond.is_cuda):
            raise RuntimeError("SLatFlowModel runs on CUDA tensors only (no CPU fallback)")
        ent = self._context_entry(cond)
        B = x.shape[0]
        key = (x.coords.data_ptr(), x.coords.shape[0], B)
        g = ent["graphs"].get(key)
        tt = t.to(self.device, F32).reshape(-1)
        if tt.numel() == 1 and B > 1:
            tt = tt.expand(B)
        if g is None:
            xs = x.feats.to(F32).clone()
            ts = tt.clone()
            sx = x.replace(xs)
            self.forward(sx, ts, cond)                         # warm-up: resampling plan, neighbour maps, workspace
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.forward(sx, ts, cond).feats
            # `sx` stays with the graph: its spatial cache owns the neighbour maps / resampling plan / APE the graph reads;
            # the last field is the token count of the workspace the graph points into (see _workspace)
            g = (graph, xs, ts, out, sx, next(reversed(self._ws)))
            if len(ent["graphs"]) >= 4:
                ent["graphs"].clear()
            ent["graphs"][key] = g
        graph, xs, ts, out, _, n_ws = g
        self._workspace(n_ws)                              # touch: keeps this graph's workspace at the young end of the LRU
        xs.copy_(x.feats)
        ts.copy_(tt)
        graph.replay()
        return x.replace(out.clone())

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x: SparseTensor, t: torch.Tensor, cond: torch.Tensor) -> SparseTensor:
        if not self._loaded:
            raise RuntimeError("load_state_dict first")
        if not (x.feats.is_cuda and cond.is_cuda):
            raise RuntimeError("SLatFlowModel runs on CUDA tensors only (no CPU fallback)")
        dev, C, B = self.device, self.model_channels, x.shape[0]
        if B > 8:
            raise ValueError("at most 8 batch entries per call")
        tt = t.to(dev, F32).reshape(-1).contiguous()
        if tt.numel() == 1 and B > 1:
            tt = tt.expand(B).contiguous()
        temb = torch.empty((B, C), dtype=F16, device=dev)
        stemb = torch.empty((B, C), dtype=F16, device=dev)
        mod = torch.empty((B, self.R), dtype=F16, device=dev)
        ops.dit_modulation(tt, self.t_w0, self.t_b0, self.t_w2, self.t_b2, self.mod_w, self.mod_b, temb, stemb, mod)
        kvs = self._context(cond)

        st = x
        h = ops.small_linear(x.feats.to(F32).contiguous(), self.in_w, self.in_b, out_f16=True)      # input_layer (:232)
        skips, levels = [], []
        for blk in self.input_blocks:                                                                # :239-242
            if blk.downsample:
                fine = st.replace(h)
                coarse = SparseDownsample(2)(fine)
                levels.append((st, downsample_plan(fine, 2)["idx"]))
                st, h = coarse, coarse.feats
            h = blk.forward(st, h, mod)
            skips.append(h)

        pos = st.get_spatial_cache(f"ape_{C}")                                                       # :244-245
        if pos is None or pos.shape[0] != h.shape[0]:
            pos = ops.ape(st.coords[:, 1:].to(F32).contiguous(), C)
            st.register_spatial_cache(f"ape_{C}", pos)
        X = torch.add(pos, h)                                                                        # fp32 residual stream
        ws = self._workspace(X.shape[0])
        for blk, kv in zip(self.blocks, kvs):                                                        # :246-247
            blk.forward(X, st.layout, mod, kv, ws)
        h = ops.cast_f16(X)

        for blk in self.out_blocks:                                                                  # :250-256
            if self.use_skip_connection:
                h = ops.gather_concat(a=h, b=skips.pop())
            if blk.upsample:
                st, idx = levels.pop()
                h = blk.forward(st, h, mod, idx=idx)
            else:
                h = blk.forward(st, h, mod)

        a = ops.ln_mod(h, eps=1e-5)                                                                  # F.layer_norm (:258)
        out = torch.empty((h.shape[0], self.out_channels), dtype=F32, device=dev)
        ops.gemm(a, self.out_w, self.out_b, ops.EPI_F32_COMPACT, out=out)                            # out_layer (:259)
        return x.replace(out)

    def __call__(self, x, t, cond):
        return self.forward_graphed(x, t, cond) if self.use_graphs else self.forward(x, t, cond)
