"""Decode engine host side: a GPTQ LLaMA decoded token by token on a static KV cache, the whole step
captured once in a CUDA graph (gptq_llama_decode_step in include/gptq_b200.h).

This is the B200-native replacement for the reference's per-token loop (llama.py:419-433 /
model.generate in llama_inference.py:120): ~480 Python-dispatched launches and an O(n) torch.cat of
the KV cache per token become one graph replay.  torch is used for device memory, streams and graph
capture only.
"""
import ctypes
import math

import torch

from . import ops
from ._lib import LlamaLayer, LlamaModel, LlamaState, LlamaTP, QWeight, check, lib

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


LLAMA_SHAPES = {  # hidden, intermediate, layers, heads
    '7b': (4096, 11008, 32, 32),
    '13b': (5120, 13824, 40, 40),
    '33b': (6656, 17920, 60, 52),
    '65b': (8192, 22016, 80, 64),
    'tiny': (256, 704, 2, 2),    # intermediate not a multiple of 256: exercises the kernel-chain engine
    'tiny256': (256, 768, 2, 2),  # eligible for the persistent single-kernel path
    'tiny512': (512, 1024, 2, 4),  # shardable over 2 tensor-parallel ranks (2 heads and 2 slabs of 256 MLP columns each)
}


class QLayerWeights:
    """Packed tensors of one QuantLinear (kept alive by the engine) + the act-order probe result."""

    def __init__(self, qweight, scales, qzeros, g_idx, bits, groupsize):
        self.qweight, self.scales, self.qzeros, self.g_idx, self.bits = qweight.contiguous(), scales.contiguous(), qzeros.contiguous(), g_idx.contiguous(), bits
        K = qweight.shape[0] * 32 // bits
        self.g_idx = self.g_idx[:K].contiguous()
        self.groupsize = groupsize
        self.hint = groupsize if ops.is_trivial_g_idx(self.g_idx, groupsize) else 0

    def kernel_form(self, allow_perm=True):
        """(layer in the layout the tuned int4 kernels take, input gather or None); see ops.kernel_form."""
        plan = ops.kernel_form(self.qweight, self.scales, self.qzeros, self.g_idx, self.bits, self.groupsize, allow_perm=allow_perm)
        if plan is None:
            return self, None
        return QLayerWeights(plan['qweight'], self.scales, plan['qzeros'], plan['g_idx'], plan['bits'], self.groupsize), plan['perm']

    def permute_columns(self, perm):
        """The same layer with output columns reordered: out'[:, j] = out[:, perm[j]] (used to fold the NEXT layer's input gather)."""
        zeros = ops.unpack_qzeros(self.qzeros, self.bits).index_select(1, perm)
        return QLayerWeights(self.qweight.index_select(1, perm), self.scales.index_select(1, perm), ops.pack_qzeros(zeros, self.bits), self.g_idx, self.bits,
                             self.groupsize)

    def column_slice(self, cols):
        """The layer restricted to output columns `cols` (a LongTensor of whole groups of 8 consecutive columns): out[:, j] = full[:, cols[j]]."""
        ipb = 32 // self.bits
        assert cols.numel() % ipb == 0 and bool((cols.view(-1, ipb)[:, 0] % ipb == 0).all()), 'column shards must keep the packed zero words whole'
        zcols = (cols.view(-1, ipb)[:, 0] // ipb).contiguous()
        return QLayerWeights(self.qweight.index_select(1, cols), self.scales.index_select(1, cols), self.qzeros.index_select(1, zcols), self.g_idx, self.bits, self.groupsize)

    def row_slice(self, k0, k1):
        """The layer restricted to input features [k0, k1) (whole quantisation groups, no act-order): a K-shard whose partial outputs add up."""
        assert self.hint == self.groupsize and k0 % self.groupsize == 0 and k1 % self.groupsize == 0 and self.bits in (2, 4, 8)
        ipb, gs = 32 // self.bits, self.groupsize
        g = (torch.arange(k1 - k0, device=self.qweight.device) // gs).to(torch.int32)
        return QLayerWeights(self.qweight[k0 // ipb:k1 // ipb], self.scales[k0 // gs:k1 // gs], self.qzeros[k0 // gs:k1 // gs], g, self.bits, gs)

    @classmethod
    def from_module(cls, m):
        return cls(m.qweight, m.scales, m.qzeros, m.g_idx, m.bits, m.groupsize)

    def struct(self) -> QWeight:
        return ops.make_qweight(self.qweight, self.scales, self.qzeros, self.g_idx, self.bits, self.hint)


def random_qlayer(K, N, bits, groupsize, device, gen, act_order=False):
    """Synthetic packed layer (SURVEY.md 8(d) perf fixture): uniform random fields, scales ~ U(1e-3, 1.1e-2)."""
    G = math.ceil(K / groupsize)
    half = 1 << (bits - 1)
    if bits == 3:
        qw = ops.pack_qweight(torch.randint(0, 8, (K, N), device=device, generator=gen, dtype=torch.int32), 3)
    else:
        qw = torch.randint(-2**31, 2**31 - 1, (K // 32 * bits, N), device=device, generator=gen, dtype=torch.int32)
    # zero points centred on the weight grid like a real asymmetric GPTQ checkpoint (stored minus one: z = stored + 1 has the mean
    # of the uniform fields, 2^(bits-1) - 1/2).  Fully random zeros give every weight the same mean offset, and a 32-layer stack of
    # such matrices amplifies the common mode of the activations until fp16 overflows.
    lo, hi = (half - 3, half + 1) if bits >= 4 else (half - 2, half)
    qz = ops.pack_qzeros(torch.randint(lo, hi, (G, N), device=device, generator=gen, dtype=torch.int32), bits)
    s = (torch.rand(G, N, device=device, generator=gen) * 1e-2 + 1e-3).half()
    g = (torch.arange(K, device=device) // groupsize).to(torch.int32)
    if act_order:
        perm = torch.randperm(K, device=device, generator=gen)
        g = g[torch.argsort(perm)].contiguous()
    return QLayerWeights(qw, s, qz, g, bits, groupsize)


def kernel_layers(layers, allow_perm=True):
    """Load-time preparation of a layer stack for the decode kernels (the stored tensors stay as they are):
    2/3-bit fields widened to nibbles and act-order rows regrouped (ops.kernel_form).  The input gathers this needs are
    returned per layer for the kernel (qkv, o, gate|up); down_proj's gather costs nothing at run time: it is folded into
    the column order of gate|up, whose SwiGLU output then comes out in down_proj's regrouped order.
    Returns (prepared layers, perms) or None when gate and up do not share their act-order map."""
    out, perms = [], []
    for ly in layers:
        k, pm = {}, {}
        for name in ('qkv', 'o', 'gate', 'up', 'down'):
            k[name], pm[name] = ly[name].kernel_form(allow_perm)
        if (pm['gate'] is None) != (pm['up'] is None) or (pm['gate'] is not None and not torch.equal(pm['gate'], pm['up'])):
            return None
        if pm['down'] is not None:
            k['gate'], k['up'] = k['gate'].permute_columns(pm['down']), k['up'].permute_columns(pm['down'])
        # per-input-feature vectors of a regrouped matvec are handed over in regrouped order
        k['input_norm'] = ly['input_norm'] if pm['qkv'] is None else ly['input_norm'].index_select(0, pm['qkv']).contiguous()
        k['post_norm'] = ly['post_norm'] if pm['gate'] is None else ly['post_norm'].index_select(0, pm['gate']).contiguous()
        out.append(k)
        perms.append({n: (pm[n].to(torch.int32).contiguous() if pm[n] is not None else None) for n in ('qkv', 'o', 'gate')})
    return out, perms


class LlamaDecoder:
    """Owns the weights, the KV cache and the captured graph; `step()` decodes one token per sequence."""

    def __init__(self, layers, embed, final_norm, lm_head, n_heads, rms_eps=1e-6, rope_base=10000.0, batch=1, max_seq=2048, use_graph=True, head_dim=None,
                 tp=None):
        """tp = (rank, size, vocab_begin, vocab_end[, reduce_mode]): this process holds one tensor-parallel shard (see shard_for_rank / gptq_llama_tp): `layers`,
        `lm_head` and `n_heads` are the LOCAL ones, head_dim must be given, and torch.distributed must be initialised (the scratch and logits
        buffers of the ranks are exchanged as CUDA IPC handles)."""
        self.dev = embed.device
        self.tp = tp
        self.head_dim = head_dim or embed.shape[1] // n_heads
        self.layers = layers  # list of dicts: qkv, o, gate, up, down (QLayerWeights), input_norm, post_norm (fp16 tensors)
        self.embed, self.final_norm, self.lm_head = embed.contiguous(), final_norm.contiguous(), lm_head.contiguous()
        self.hidden = embed.shape[1]
        self.vocab = embed.shape[0]
        self.n_heads = n_heads
        self.intermediate = layers[0]['gate'].qweight.shape[1]
        self.batch, self.max_seq = batch, max_seq
        with torch.cuda.device(self.dev):
            # kernel-side view of the weights: nibble-widened 2/3-bit fields, regrouped act-order rows (+ input gathers)
            prepared = kernel_layers(layers, allow_perm=(batch == 1))
            if prepared is None:
                prepared = kernel_layers(layers, allow_perm=False)
            self.klayers, self.perms = prepared
            self._layer_arr = (LlamaLayer * len(layers))()
            self._fill_layer_structs()
            m = LlamaModel()
            m.n_layers, m.hidden, m.n_heads, m.head_dim = len(layers), self.hidden, n_heads, self.head_dim
            m.intermediate, m.vocab, m.rms_eps, m.rope_base = self.intermediate, self.vocab, rms_eps, rope_base
            m.layers = ctypes.cast(self._layer_arr, ctypes.POINTER(LlamaLayer))
            m.embed, m.final_norm, m.lm_head = self.embed.data_ptr(), self.final_norm.data_ptr(), self.lm_head.data_ptr()
            self.model = m
            cache_shape = (len(layers), batch, n_heads, max_seq, self.head_dim)
            self.k_cache = torch.zeros(cache_shape, dtype=torch.float16, device=self.dev)
            self.v_cache = torch.zeros(cache_shape, dtype=torch.float16, device=self.dev)
            self.tokens = torch.zeros(batch, dtype=torch.int32, device=self.dev)
            self.positions = torch.zeros(batch, dtype=torch.int32, device=self.dev)
            self.next_tokens = torch.zeros(batch, dtype=torch.int32, device=self.dev)
            nbytes = lib.gptq_llama_scratch_bytes(ctypes.byref(m), batch, max_seq)
            if nbytes == 0:
                raise ValueError('unsupported decode configuration (batch must be 1..8)')
            self._tp_struct = None
            if tp is None:
                self.logits = torch.zeros(batch, self.vocab, dtype=torch.float16, device=self.dev)
                self.scratch = torch.zeros(nbytes, dtype=torch.uint8, device=self.dev)
            else:
                self._setup_tp(nbytes)
            st = LlamaState()
            st.batch, st.max_seq = batch, max_seq
            st.k_cache, st.v_cache = self.k_cache.data_ptr(), self.v_cache.data_ptr()
            st.tokens, st.positions = self.tokens.data_ptr(), self.positions.data_ptr()
            st.logits, st.next_tokens = self.logits.data_ptr(), self.next_tokens.data_ptr()
            st.scratch, st.scratch_bytes = self.scratch.data_ptr(), nbytes
            if self._tp_struct is not None:
                st.tp = ctypes.pointer(self._tp_struct)
            self.state = st
            if any(p is not None for pm in self.perms for p in pm.values()) and self.launches_per_step() != 1:
                # the input gathers exist only in the persistent kernel: act-order layers go back to their stored form
                self.klayers, self.perms = kernel_layers(layers, allow_perm=False)
                self._fill_layer_structs()
        self.n_launches = None
        self.graph = None
        self._stream = torch.cuda.Stream(self.dev)
        if use_graph:
            self._capture()

    def _setup_tp(self, scratch_bytes):
        """Scratch and logits in IPC-shareable device memory; every rank maps every other rank's buffers (gptq_llama_tp)."""
        import torch.distributed as dist
        rank, size, v0, v1 = self.tp[:4]
        assert self.batch == 1 and dist.is_initialized() and dist.get_world_size() == size

        def ipc_buffer(nbytes, dtype):
            ptr, handle = ctypes.c_void_p(), ctypes.create_string_buffer(64)
            check(lib.gptq_ipc_alloc(nbytes, ctypes.byref(ptr), handle))

            class _Raw:  # zero-copy torch view of the raw allocation
                __cuda_array_interface__ = {'shape': (nbytes, ), 'typestr': '|u1', 'data': (ptr.value, False), 'version': 2}

            t = torch.as_tensor(_Raw(), device=self.dev)
            return ptr.value, handle.raw, t.view(dtype)

        self._scratch_ptr, h_scr, self.scratch = ipc_buffer(scratch_bytes, torch.uint8)
        self._logits_ptr, h_log, logits = ipc_buffer(self.batch * self.vocab * 2, torch.float16)
        self.logits = logits.view(self.batch, self.vocab)
        handles = [None] * size
        dist.all_gather_object(handles, (h_scr, h_log))
        t = LlamaTP()
        t.size, t.rank, t.vocab_begin, t.vocab_end = size, rank, v0, v1
        t.reduce_mode = self.tp[4] if len(self.tp) > 4 else 0
        self._peer_maps = []
        for q, (hs, hl) in enumerate(handles):
            if q == rank:
                t.peer_scratch[q], t.peer_logits[q] = self._scratch_ptr, self._logits_ptr
                continue
            for field, h in ((t.peer_scratch, hs), (t.peer_logits, hl)):
                ptr = ctypes.c_void_p()
                check(lib.gptq_ipc_open(h, ctypes.byref(ptr)))
                field[q] = ptr.value
                self._peer_maps.append(ptr.value)
        self._tp_struct = t
        dist.barrier()  # every rank has mapped every buffer before anybody launches

    def _fill_layer_structs(self):
        for i, ly in enumerate(self.klayers):
            for name in ('qkv', 'o', 'gate', 'up', 'down'):
                setattr(self._layer_arr[i], name, ly[name].struct())
            self._layer_arr[i].input_norm = ly['input_norm'].data_ptr()
            self._layer_arr[i].post_norm = ly['post_norm'].data_ptr()
            pm = self.perms[i]
            self._layer_arr[i].qkv_perm = pm['qkv'].data_ptr() if pm['qkv'] is not None else None
            self._layer_arr[i].o_perm = pm['o'].data_ptr() if pm['o'] is not None else None
            self._layer_arr[i].mlp_perm = pm['gate'].data_ptr() if pm['gate'] is not None else None

    # ------------------------------------------------------------------------------------------
    def _enqueue(self, stream):
        check(lib.gptq_llama_decode_step(ctypes.byref(self.model), ctypes.byref(self.state), ctypes.c_void_p(stream.cuda_stream)))

    def _capture(self):
        with torch.cuda.device(self.dev):
            torch.cuda.synchronize()
            with torch.cuda.stream(self._stream):
                self._enqueue(self._stream)  # warm-up outside capture (lazy module loading, func attributes)
                self._stream.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._stream):
                    self._enqueue(self._stream)
            self.graph = g
            torch.cuda.synchronize()

    def launches_per_step(self) -> int:
        """Kernels of ours launched per decoded token (1 on the persistent single-kernel path)."""
        return int(lib.gptq_llama_decode_launches(ctypes.byref(self.model), ctypes.byref(self.state)))

    def mega_scratch_offset(self) -> int:
        """Byte offset of the persistent kernel's region in self.scratch (diagnostics; see gptq_llama_persistent_scratch_offset)."""
        return int(lib.gptq_llama_persistent_scratch_offset(ctypes.byref(self.model), self.batch, self.max_seq))

    def step(self, stream=None):
        """Run one decode step on the tokens/positions currently in device memory."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._enqueue(stream or torch.cuda.current_stream(self.dev))

    def reset(self):
        self.positions.zero_()

    def set_input(self, tokens, position):
        """Token ids (int or sequence of `batch` ints) and the cache position of this step; validated on the host: the kernels
        only clamp (a position beyond the cache or a token outside the vocabulary must never reach them)."""
        toks = [int(tokens)] * self.batch if isinstance(tokens, int) else [int(t) for t in tokens]
        if len(toks) != self.batch:
            raise ValueError(f'expected {self.batch} token ids, got {len(toks)}')
        if not 0 <= int(position) < self.max_seq:
            raise ValueError(f'position {position} outside the KV cache (max_seq = {self.max_seq})')
        if any(t < 0 or t >= self.vocab for t in toks):
            raise ValueError(f'token id outside the vocabulary (0..{self.vocab - 1})')
        self.tokens.copy_(torch.tensor(toks, dtype=torch.int32))
        self.positions.fill_(int(position))

    @torch.no_grad()
    def prefill(self, prompt_ids):
        """Batched pass over the first len(prompt) - 1 prompt tokens that fills the static KV cache (batch 1): per layer the quantized linears on
        the tcgen05 GEMM path (gptq_qlinear_fwd / gptq_fused_mlp_fwd with M = tokens), the RoPE and RMSNorm kernels, torch SDPA for the causal
        attention (as the reference's QuantLlamaAttention does, quant/fused_attn.py:154-155); keys are cached after RoPE.  The LAST prompt token then
        goes through the decode step like every generated one (it produces the first logits).  Returns the number of cached positions."""
        assert self.batch == 1 and self.tp is None
        n = len(prompt_ids) - 1
        if n <= 0:
            return 0
        H, nh, hd = self.hidden, self.n_heads, self.head_dim
        w4 = lambda w: (w.qweight, w.scales, w.qzeros, w.g_idx)
        x = self.embed[torch.tensor(list(prompt_ids[:n]), device=self.dev)]  # [n, H]
        pos = torch.arange(n, device=self.dev, dtype=torch.int64)[None, :]
        for li, ly in enumerate(self.layers):
            qkv = ops.matmul248(ops.rmsnorm(x, ly['input_norm'], self.model.rms_eps), *w4(ly['qkv']), ly['qkv'].bits, groupsize=ly['qkv'].hint).view(1, n, 3, nh, hd)
            ops.rotate_half_(qkv[:, :, :2], pos, base=self.model.rope_base)
            q, k, v = (qkv[0, :, i].transpose(0, 1) for i in range(3))  # [nh, n, hd]
            self.k_cache[li, 0, :, :n] = k
            self.v_cache[li, 0, :, :n] = v
            att = torch.nn.functional.scaled_dot_product_attention(q[None], k[None], v[None], is_causal=True)[0].transpose(0, 1).reshape(n, H)
            x = x + ops.matmul248(att, *w4(ly['o']), ly['o'].bits, groupsize=ly['o'].hint)
            h = ops.fused_mlp(ops.rmsnorm(x, ly['post_norm'], self.model.rms_eps), w4(ly['gate']), w4(ly['up']), ly['gate'].bits, ly['gate'].hint)
            x = x + ops.matmul248(h, *w4(ly['down']), ly['down'].bits, groupsize=ly['down'].hint)
        return n

    @torch.no_grad()
    def generate(self, prompt_ids, max_new_tokens, prefill=True):
        """Greedy decode (batch 1).  One engine, two phases: the prompt is prefilled in one batched pass (tcgen05 GEMM path) into the static KV
        cache, then the persistent decode kernel takes over token by token; prefill=False feeds the prompt through the decode step instead."""
        assert self.batch == 1
        if len(prompt_ids) < 1 or len(prompt_ids) + max_new_tokens > self.max_seq + 1:
            raise ValueError(f'prompt ({len(prompt_ids)}) + max_new_tokens ({max_new_tokens}) does not fit the KV cache (max_seq = {self.max_seq})')
        if any(int(t) < 0 or int(t) >= self.vocab for t in prompt_ids):
            raise ValueError(f'prompt token id outside the vocabulary (0..{self.vocab - 1})')
        out = list(prompt_ids)
        self.reset()
        start = self.prefill(prompt_ids) if (prefill and self.tp is None) else 0
        tok = torch.empty(1, dtype=torch.int32, device=self.dev)
        for i in range(start, len(prompt_ids) + max_new_tokens - 1):
            if i < len(prompt_ids):
                self.tokens.copy_(torch.tensor([prompt_ids[i]], dtype=torch.int32), non_blocking=False)
            else:
                self.tokens.copy_(tok)
            self.positions.fill_(i)
            self.step()
            tok.copy_(self.next_tokens)
            if i >= len(prompt_ids) - 1:
                out.append(int(tok.item()))
        return out


def synthetic_llama(size='7b', bits=4, groupsize=128, act_order=False, vocab=32000, device='cuda:0', seed=0, n_layers=None, **kw):
    """Random-init GPTQ LLaMA of the named size (no checkpoints are reachable offline): every layer has its
    own distinct packed tensors so that a decode step streams the full model from HBM."""
    hidden, inter, layers, heads = LLAMA_SHAPES[size]
    layers = n_layers or layers
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(seed)
    L = []
    for _ in range(layers):
        gate = random_qlayer(hidden, inter, bits, groupsize, dev, gen, act_order)
        up = random_qlayer(hidden, inter, bits, groupsize, dev, gen, act_order)
        if act_order:  # gate and up see the same input, hence the same Hessian diagonal and the same act-order map (gptq.py:210-216)
            up = QLayerWeights(up.qweight, up.scales, up.qzeros, gate.g_idx.clone(), bits, groupsize)
        L.append(
            dict(qkv=random_qlayer(hidden, 3 * hidden, bits, groupsize, dev, gen, act_order), o=random_qlayer(hidden, hidden, bits, groupsize, dev, gen, act_order),
                 gate=gate, up=up, down=random_qlayer(inter, hidden, bits, groupsize, dev, gen, act_order),
                 input_norm=(torch.rand(hidden, device=dev, generator=gen) * 0.2 + 0.9).half(),
                 post_norm=(torch.rand(hidden, device=dev, generator=gen) * 0.2 + 0.9).half()))
    # q/k/v share their input, hence their act-order map (quant/fused_attn.py:180): nothing to do, qkv is one layer here
    embed = (torch.randn(vocab, hidden, device=dev, generator=gen) * 0.5).half()
    lm_head = (torch.randn(vocab, hidden, device=dev, generator=gen) * 0.02).half()
    final_norm = (torch.rand(hidden, device=dev, generator=gen) * 0.2 + 0.9).half()
    return LlamaDecoder(L, embed, final_norm, lm_head, heads, **kw)


def shard_for_rank(layers, lm_head, n_heads, head_dim, rank, size):
    """Tensor-parallel shard of a layer stack (BASELINE config 5): qkv columns of this rank's heads (q | k | v), o_proj rows
    of the same heads, gate/up column slabs of 256 with the matching down_proj rows, lm_head rows.  Plain g_idx only (row shards keep whole groups).
    Returns (local layers, local lm_head, local heads, (vocab_begin, vocab_end))."""
    assert n_heads % size == 0, 'heads must divide over the ranks'
    hl = n_heads // size
    H = n_heads * head_dim
    dev = lm_head.device
    hcols = torch.arange(rank * hl * head_dim, (rank + 1) * hl * head_dim, device=dev)
    out = []
    for ly in layers:
        inter = ly['gate'].qweight.shape[1]
        nslab = inter // 256
        c0, c1 = rank * nslab // size * 256, (rank + 1) * nslab // size * 256
        mcols = torch.arange(c0, c1, device=dev)
        out.append(dict(qkv=ly['qkv'].column_slice(torch.cat([hcols, H + hcols, 2 * H + hcols])), o=ly['o'].row_slice(int(hcols[0]), int(hcols[-1]) + 1),
                        gate=ly['gate'].column_slice(mcols), up=ly['up'].column_slice(mcols), down=ly['down'].row_slice(c0, c1), input_norm=ly['input_norm'],
                        post_norm=ly['post_norm']))
    V = lm_head.shape[0]
    v0, v1 = rank * V // size, (rank + 1) * V // size
    return out, lm_head[v0:v1].contiguous(), hl, (v0, v1)


def synthetic_llama_tp(size_name, rank, world, bits=4, groupsize=128, vocab=32000, device='cuda:0', seed=0, n_layers=None, full=None, reduce_mode=0, **kw):
    """One tensor-parallel rank of the random-init model `synthetic_llama(size_name, seed=seed)` (every rank generates the same full model from the same
    seed layer by layer and keeps its shard), or of the given `full` decoder's weights."""
    hidden, inter, layers, heads = LLAMA_SHAPES[size_name]
    layers = n_layers or layers
    dev = torch.device(device)
    hd = hidden // heads
    if full is not None:
        L, lm_head, hl, (v0, v1) = shard_for_rank(full.layers, full.lm_head, heads, hd, rank, world)
        return LlamaDecoder(L, full.embed, full.final_norm, lm_head, hl, head_dim=hd, tp=(rank, world, v0, v1, reduce_mode), **kw)
    gen = torch.Generator(device=dev).manual_seed(seed)
    L = []
    fake_head = torch.empty(0, hidden, device=dev)
    for _ in range(layers):  # shard as we go: a 65B model never exists whole on one GPU
        gate = random_qlayer(hidden, inter, bits, groupsize, dev, gen)
        up = random_qlayer(hidden, inter, bits, groupsize, dev, gen)
        ly = dict(qkv=random_qlayer(hidden, 3 * hidden, bits, groupsize, dev, gen), o=random_qlayer(hidden, hidden, bits, groupsize, dev, gen), gate=gate, up=up,
                  down=random_qlayer(inter, hidden, bits, groupsize, dev, gen), input_norm=(torch.rand(hidden, device=dev, generator=gen) * 0.2 + 0.9).half(),
                  post_norm=(torch.rand(hidden, device=dev, generator=gen) * 0.2 + 0.9).half())
        L.append(shard_for_rank([ly], fake_head, heads, hd, rank, world)[0][0])
        del ly, gate, up
    embed = (torch.randn(vocab, hidden, device=dev, generator=gen) * 0.5).half()
    lm_head = (torch.randn(vocab, hidden, device=dev, generator=gen) * 0.02).half()
    final_norm = (torch.rand(hidden, device=dev, generator=gen) * 0.2 + 0.9).half()
    v0, v1 = rank * vocab // world, (rank + 1) * vocab // world
    return LlamaDecoder(L, embed, final_norm, lm_head[v0:v1].contiguous(), heads // world, head_dim=hd, tp=(rank, world, v0, v1, reduce_mode), **kw)


def from_hf_quant_model(model, batch=1, max_seq=2048, **kw):
    """Build a decoder from an HF LlamaForCausalLM that went through the reference's load_quant recipe with this
    repo's quant package (make_quant_linear -> load_state_dict -> make_quant_attn / make_quant_norm / make_fused_mlp)."""
    import quant
    L = []
    for layer in model.model.layers:
        attn, mlp = layer.self_attn, layer.mlp
        if not isinstance(attn, quant.QuantLlamaAttention):
            raise ValueError('call quant.make_quant_attn(model) first')
        if isinstance(mlp, quant.QuantLlamaMLP):
            gate = QLayerWeights(mlp.gate_proj_qweight, mlp.gate_proj_scales, mlp.gate_proj_qzeros, mlp.gate_proj_g_idx, mlp.bits, mlp.groupsize)
            up = QLayerWeights(mlp.up_proj_qweight, mlp.up_proj_scales, mlp.up_proj_qzeros, mlp.up_proj_g_idx, mlp.bits, mlp.groupsize)
        else:
            gate, up = QLayerWeights.from_module(mlp.gate_proj), QLayerWeights.from_module(mlp.up_proj)
        L.append(
            dict(qkv=QLayerWeights.from_module(attn.qkv_proj), o=QLayerWeights.from_module(attn.o_proj), gate=gate, up=up, down=QLayerWeights.from_module(mlp.down_proj),
                 input_norm=layer.input_layernorm.weight.data.half().contiguous(), post_norm=layer.post_attention_layernorm.weight.data.half().contiguous()))
    cfg = model.config
    return LlamaDecoder(L, model.model.embed_tokens.weight.data.half(), model.model.norm.weight.data.half().contiguous(), model.lm_head.weight.data.half(),
                        cfg.num_attention_heads, rms_eps=cfg.rms_norm_eps, rope_base=getattr(cfg, 'rope_theta', 10000.0) or 10000.0, batch=batch, max_seq=max_seq, **kw)
