"""Model-level engines over the C-ABI handle (include/dimb200.h, "Model level").

VQEngine   : VQAutoEncoder.encode / decode   (/root/reference/code/models/stage1_BIWI.py:22-37)
SLMFTEngine: SLMFT.forward_encoder + context, decoder_joint.generate (seq2seq_pretrain.py:431-452)

Weights are registered by their reference state_dict keys; the engines keep the CUDA tensors alive (the library
borrows the pointers).  Workspaces are torch allocations owned here and reused across calls.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .schema import S2SConfig, VQConfig

PREC_FP32, PREC_BF16, PREC_FP32_TC = 0, 1, 2


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _index_arg(t, B, device, name, lo, hi):
    """Normalise an optional per-clip int32 argument of the C ABI (`lens`, `batch_index`): the library reads DEVICE int32 memory,
    so an int64 or host tensor must never reach it as a raw pointer.  Shape (B,) is enforced; values are range-checked here when
    the tensor lives on the host (no synchronisation) -- device tensors are clamped into range by the kernels instead
    (lens to [1, T], batch_index to [0, pe_max_len)), so a bad value can never read outside the buffers."""
    if t is None:
        return None
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if t.shape != (B,):
        raise ValueError(f"{name} must have shape ({B},), got {tuple(t.shape)}")
    if t.is_floating_point() or t.dtype == torch.bool:
        raise ValueError(f"{name} must be an integer tensor")
    if not t.is_cuda and B > 0:
        mn, mx = int(t.min()), int(t.max())
        if mn < lo or mx >= hi:
            raise ValueError(f"{name} values must lie in [{lo}, {hi}), got [{mn}, {mx}]")
    return t.to(device=device, dtype=torch.int32).contiguous()


class Handle:
    """dim_handle_t with the registered tensors kept alive."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("dim_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        h = C.c_void_p()
        _lib.check(self.lib.dim_create(C.byref(h), self.device.index), "dim_create")
        self.h = h
        self.keep = {}

    def register(self, state_dict, prefix=""):
        """Copy (if needed) every fp32 tensor of `state_dict` to the device and register it under prefix+key."""
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not v.is_floating_point():
                continue
            t = v.detach().to(device=self.device, dtype=torch.float32).contiguous()
            name = prefix + k
            self.keep[name] = t
            shape = (C.c_int64 * t.dim())(*t.shape)
            _lib.check(self.lib.dim_set_tensor(self.h, name.encode(), t.data_ptr(), 0, t.dim(), shape), "dim_set_tensor")

    def close(self):
        if getattr(self, "h", None):
            torch.cuda.synchronize(self.device)
            self.lib.dim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _Workspace:
    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self.buf


class VQEngine:
    """One VQ-VAE built from the tensors registered under `prefix`.  `encoder` / `decoder` name the two halves (None: absent);
    `out_dim` is the decoder's output width (default in_dim).  VQSpeakerAutoEncoder (stage1_BIWI.py:140-173) is three engines over
    one codebook: ("encoder", None), (None, "decoder_v", out_dim=56), (None, "decoder_a", out_dim=768)."""

    def __init__(self, handle: Handle, cfg: VQConfig = VQConfig(), prefix: str = "", precision: int = PREC_FP32,
                 encoder: str | None = "encoder", decoder: str | None = "decoder", out_dim: int | None = None):
        self.handle, self.cfg, self.prefix, self.precision = handle, cfg, prefix, precision
        self.fqn = max(1, cfg.face_quan_num)
        self.out_dim = out_dim or cfg.in_dim
        cc = _lib.VQConfigC(cfg.in_dim, cfg.hidden_size, cfg.num_hidden_layers, cfg.num_attention_heads,
                            cfg.intermediate_size, cfg.n_embed, cfg.zquant_dim, cfg.pe_max_len, cfg.neg, self.fqn, self.out_dim)
        torch.cuda.synchronize(handle.device)
        m = C.c_int(-1)
        _lib.check(handle.lib.dim_vqvae_build_parts(handle.h, prefix.encode(), encoder.encode() if encoder else None,
                                                    decoder.encode() if decoder else None, C.byref(cc), precision, C.byref(m)),
                   "dim_vqvae_build_parts")
        self.model = m.value
        self.ws = _Workspace(handle.device)

    def _workspace(self, B, T):
        n = self.handle.lib.dim_vqvae_workspace_bytes(self.handle.h, self.model, B, T)
        return self.ws.get(n), n

    def encode(self, x, lens=None, batch_index=None, want_z=False, want_quant=False):
        """x (B,T,in_dim) fp32 cuda -> idx (B,T*fqn) int64 [, z (B,T,fqn*zdim), quant (B,zdim,T*fqn)]."""
        assert x.is_cuda and x.dtype == torch.float32
        x = x.contiguous()
        B, T, _ = x.shape
        dev = x.device
        lens = _index_arg(lens, B, dev, "lens", 1, T + 1)            # an empty clip has no replicate-padding source (reference raises too)
        batch_index = _index_arg(batch_index, B, dev, "batch_index", 0, self.cfg.pe_max_len)
        idx = torch.empty(B, T * self.fqn, dtype=torch.int64, device=dev)
        z = torch.empty(B, T, self.fqn * self.cfg.zquant_dim, dtype=torch.float32, device=dev) if want_z else None
        q = torch.empty(B, self.cfg.zquant_dim, T * self.fqn, dtype=torch.float32, device=dev) if want_quant else None
        ws, n = self._workspace(B, T)
        _lib.check(self.handle.lib.dim_vqvae_encode(self.handle.h, self.model, x.data_ptr(), _ptr(lens), _ptr(batch_index),
                                                    B, T, idx.data_ptr(), _ptr(z), _ptr(q), ws.data_ptr(), n, _stream()),
                   "dim_vqvae_encode")
        return idx, z, q

    def decode(self, codes=None, quant=None, batch_index=None):
        """codes (B,L*fqn) int64 or quant (B,zdim,L*fqn) fp32 -> frames (B,L,out_dim)."""
        src = codes if codes is not None else quant
        assert src.is_cuda
        if codes is not None:
            codes = codes.contiguous()
            B, L = codes.shape
        else:
            quant = quant.contiguous()
            B, _, L = quant.shape
        assert L % self.fqn == 0, "the code sequence must hold face_quan_num codes per frame"
        L //= self.fqn
        batch_index = _index_arg(batch_index, B, src.device, "batch_index", 0, self.cfg.pe_max_len)
        out = torch.empty(B, L, self.out_dim, dtype=torch.float32, device=src.device)
        ws, n = self._workspace(B, L)
        _lib.check(self.handle.lib.dim_vqvae_decode(self.handle.h, self.model, _ptr(codes), _ptr(quant), _ptr(batch_index),
                                                    B, L, out.data_ptr(), ws.data_ptr(), n, _stream()), "dim_vqvae_decode")
        return out


class SLMFTEngine:
    def __init__(self, handle: Handle, cfg: S2SConfig = S2SConfig(), precision: int = PREC_FP32):
        self.handle, self.cfg, self.precision = handle, cfg, precision
        cc = _lib.S2SConfigC(cfg.dim_in, cfg.dim, cfg.dim_audio, cfg.depth, cfg.heads, cfg.dim_head, cfg.max_seq_len,
                             cfg.num_tokens, cfg.ff_mult)
        torch.cuda.synchronize(handle.device)
        m = C.c_int(-1)
        _lib.check(handle.lib.dim_slmft_build(handle.h, C.byref(cc), precision, C.byref(m)), "dim_slmft_build")
        self.model = m.value
        self.ws = _Workspace(handle.device)

    def _workspace(self, B, T, steps):
        n = self.handle.lib.dim_slmft_workspace_bytes(self.handle.h, self.model, B, T, steps)
        return self.ws.get(n), n

    def context(self, v_speaker, v_audio, mask, want="ctx"):
        """want="ctx": generate's context (B,T,dim+dim_audio); want="x_s": forward_encoder's output (B,T,dim)
        (seq2seq_pretrain.py:431-446)."""
        v_speaker = v_speaker.contiguous()
        B, T, _ = v_speaker.shape
        m8 = mask.to(torch.uint8).contiguous()
        dev = v_speaker.device
        ctx = xs = None
        if want == "ctx":
            v_audio = v_audio.contiguous()
            ctx = torch.empty(B, T, self.cfg.dec_dim, dtype=torch.float32, device=dev)
        else:
            xs = torch.empty(B, T, self.cfg.dim, dtype=torch.float32, device=dev)
        ws, n = self._workspace(B, T, 0)
        _lib.check(self.handle.lib.dim_slmft_context(self.handle.h, self.model, v_speaker.data_ptr(), _ptr(v_audio) if ctx is not None else None,
                                                     m8.data_ptr(), B, T, _ptr(ctx), _ptr(xs), ws.data_ptr(), n, _stream()),
                   "dim_slmft_context")
        return ctx if want == "ctx" else xs

    ENCODERS = {"encoder_s": 0, "encoder_l": 1, "encoder_joint": 2}
    NORMS = {None: 0, "norm_s": 1, "norm_l": 2, "norm": 3}

    def encode(self, which, x, mask, causal=False, norm=None, add=None):
        """One ContinuousTransformerWrapper call (return_embeddings=True) of the SLM forward (seq2seq_pretrain.py:216-224):
        encoder `which` over x (B,T,dim_in of that encoder) under the key-padding mask [and the causal attn_mask], then the
        LayerNorm head `norm` (None, "norm_s", "norm_l", "norm").  add: (dim_in,) row added to every frame (a patch embedding)."""
        x = x.float().contiguous()
        B, T, _ = x.shape
        m8 = mask.to(torch.uint8).contiguous()
        a = None if add is None else add.detach().float().reshape(-1).contiguous()
        out = torch.empty(B, T, self.cfg.dim, dtype=torch.float32, device=x.device)
        ws, n = self._workspace(B, T, 0)
        _lib.check(self.handle.lib.dim_slmft_encode(self.handle.h, self.model, self.ENCODERS[which], x.data_ptr(), _ptr(a), m8.data_ptr(),
                                                    int(bool(causal)), self.NORMS[norm], B, T, out.data_ptr(), ws.data_ptr(), n, _stream()),
                   "dim_slmft_encode")
        return out

    def teacher_forced(self, ctx, mask, tokens, kv_mask=None):
        """Teacher-forced decoder forward (seq2seq_pretrain.py:447-448): tokens (B,L) int64 decoder inputs (pad already substituted
        for ignore_index), kv_mask (B,L) bool/uint8 or None -> logits (B,L,num_tokens) fp32.  Forward only."""
        B, T, _ = ctx.shape
        L = tokens.shape[1]
        ctx = ctx.contiguous()
        m8 = mask.to(torch.uint8).contiguous()
        tokens = tokens.to(torch.int64).contiguous()
        k8 = None if kv_mask is None else kv_mask.to(torch.uint8).contiguous()
        logits = torch.empty(B, L, self.cfg.num_tokens, dtype=torch.float32, device=ctx.device)
        n = self.handle.lib.dim_slmft_teacher_forced_workspace_bytes(self.handle.h, self.model, B, T, L)
        ws = self.ws.get(n)
        _lib.check(self.handle.lib.dim_slmft_teacher_forced(self.handle.h, self.model, ctx.data_ptr(), m8.data_ptr(), tokens.data_ptr(),
                                                            _ptr(k8), B, T, L, logits.data_ptr(), ws.data_ptr(), n, _stream()),
                   "dim_slmft_teacher_forced")
        return logits

    def generate_samples(self, ctx, mask, prompt, steps, samples, uniforms, temperature=1.0, top_k=None, return_logits=False):
        """`samples` independent draws per clip over one projection of its context (x_engine_pt.py:255-270's best-of-N loop):
        uniforms (B,samples,steps) -> codes (B,samples,steps) int64 [, logits (B,samples,steps,V)]."""
        B, T, _ = ctx.shape
        ctx = ctx.contiguous()
        m8 = mask.to(torch.uint8).contiguous()
        prompt = prompt.reshape(B).contiguous()
        if top_k is None:
            top_k = math.ceil(self.cfg.top_k_frac * self.cfg.num_tokens)
        uniforms = uniforms.to(device=ctx.device, dtype=torch.float32).contiguous()
        assert uniforms.shape == (B, samples, steps)
        out = torch.empty(B, samples, steps, dtype=torch.int64, device=ctx.device)
        logits = torch.empty(B, samples, steps, self.cfg.num_tokens, dtype=torch.float32, device=ctx.device) if return_logits else None
        n = self.handle.lib.dim_slmft_samples_workspace_bytes(self.handle.h, self.model, B, T, steps, samples)
        ws = self.ws.get(n)
        _lib.check(self.handle.lib.dim_slmft_generate_samples(self.handle.h, self.model, ctx.data_ptr(), m8.data_ptr(),
                                                              prompt.data_ptr(), B, T, steps, samples, float(temperature), int(top_k),
                                                              uniforms.data_ptr(), out.data_ptr(), _ptr(logits), ws.data_ptr(), n,
                                                              _stream()), "dim_slmft_generate_samples")
        return (out, logits) if return_logits else out

    def generate(self, ctx, mask, prompt, steps, temperature=0.0, top_k=None, uniforms=None, return_logits=False):
        """prompt (B,) or (B,1) int64 -> codes (B,steps) int64 (seq2seq_pretrain.py:450)."""
        B, T, _ = ctx.shape
        ctx = ctx.contiguous()
        m8 = mask.to(torch.uint8).contiguous()
        prompt = prompt.reshape(B).contiguous()
        if top_k is None:
            top_k = math.ceil(self.cfg.top_k_frac * self.cfg.num_tokens)
        out = torch.empty(B, steps, dtype=torch.int64, device=ctx.device)
        logits = torch.empty(B, steps, self.cfg.num_tokens, dtype=torch.float32, device=ctx.device) if return_logits else None
        if uniforms is not None:
            uniforms = uniforms.to(device=ctx.device, dtype=torch.float32).contiguous()
            assert uniforms.shape == (B, steps)
        ws, n = self._workspace(B, T, steps)
        _lib.check(self.handle.lib.dim_slmft_generate(self.handle.h, self.model, ctx.data_ptr(), m8.data_ptr(),
                                                      prompt.data_ptr(), B, T, steps, float(temperature), int(top_k),
                                                      _ptr(uniforms), out.data_ptr(), _ptr(logits), ws.data_ptr(), n,
                                                      _stream()), "dim_slmft_generate")
        return (out, logits) if return_logits else out
