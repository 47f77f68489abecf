"""One Python function per operator-level C-ABI entry point (include/dimb200.h).

Arguments are CUDA torch tensors; only their device pointers cross the ABI, and the call is enqueued on torch's current
stream.  Nothing here computes on the host and nothing falls back to torch ops.
"""
from __future__ import annotations

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _req(t, dtype, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous CUDA tensor of dtype {dtype}")
    return t


def vq_argmin(z: torch.Tensor, codebook: torch.Tensor) -> torch.Tensor:
    """z (N,D) fp32, codebook (K,D) fp32 -> nearest code per row, (N,) int64.  quantizer.py:38-45."""
    _req(z, torch.float32, "z"); _req(codebook, torch.float32, "codebook")
    N, D = z.shape
    K = codebook.shape[0]
    idx = torch.empty(N, dtype=torch.int64, device=z.device)
    if N:
        _lib.check(_lib.load().dim_vq_argmin(_ptr(z), _ptr(codebook), _ptr(idx), N, D, K, _stream()), "dim_vq_argmin")
    return idx


def vq_gather(idx: torch.Tensor, codebook: torch.Tensor, count_bad: bool = False):
    """idx (N,) int64 -> codebook rows (N,D) fp32.  quantizer.py:79-90 / seq2seq_pretrain.py:457-461."""
    _req(idx, torch.int64, "idx"); _req(codebook, torch.float32, "codebook")
    N = idx.numel()
    K, D = codebook.shape
    out = torch.empty(N, D, dtype=torch.float32, device=idx.device)
    bad = torch.zeros(1, dtype=torch.int32, device=idx.device) if count_bad else None
    if N:
        _lib.check(_lib.load().dim_vq_gather(_ptr(idx), _ptr(codebook), _ptr(out), N, D, K, _ptr(bad), _stream()),
                   "dim_vq_gather")
    return (out, bad) if count_bad else out


def linear(a, w, bias=None, residual=None, act=0, slope=0.0):
    """act(a @ w.T + bias) + residual, fp32.  a (M,K), w (N,K)."""
    _req(a, torch.float32, "a"); _req(w, torch.float32, "w")
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().dim_linear_f32(_ptr(a), K, _ptr(w), _ptr(bias), _ptr(residual), N, _ptr(out), N, M, N, K,
                                          act, float(slope), _stream()), "dim_linear_f32")
    return out


def linear_ragged(a, w, bias=None, act=0, slope=0.0):
    """act(a @ w.T + bias) with no alignment requirement (K or N = 70110: seq2seq_pretrain.py:777, :803-807).  a (M,K), w (N,K)."""
    _req(a, torch.float32, "a"); _req(w, torch.float32, "w")
    M, K = a.shape
    N = w.shape[0]
    lib = _lib.load()
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    nws = lib.dim_linear_ragged_workspace_bytes(M, N, K)
    ws = torch.empty(max(nws, 1), dtype=torch.uint8, device=a.device)
    _lib.check(lib.dim_linear_ragged_f32(_ptr(a), K, _ptr(w), K, _ptr(bias), _ptr(out), N, M, N, K, act, float(slope),
                                         _ptr(ws), nws, _stream()), "dim_linear_ragged_f32")
    return out


def lstm(x, params, hidden, layers=2, bidirectional=True, prefix=""):
    """nn.LSTM(batch_first=True) forward with zero initial state (seq2seq_pretrain.py:789-802): x (B,T,in) -> (B,T,ndir*hidden).
    `params`: mapping with torch's parameter names (weight_ih_l0, weight_hh_l0, bias_ih_l0, bias_hh_l0, *_reverse, ...)."""
    _req(x, torch.float32, "x")
    B, T, _ = x.shape
    lib = _lib.load()
    ndir = 2 if bidirectional else 1
    cur = x
    for k in range(layers):
        nws = lib.dim_lstm_layer_workspace_bytes(B, T, cur.shape[-1], hidden, ndir)
        ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
        names = [f"{prefix}{n}_l{k}{sfx}" for sfx in ("", "_reverse") for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
        w = [_req(params[n].detach(), torch.float32, n) for n in names[:4]]
        w += [_req(params[n].detach(), torch.float32, n) for n in names[4:]] if bidirectional else [None] * 4
        out = torch.empty(B, T, ndir * hidden, dtype=torch.float32, device=x.device)
        _lib.check(lib.dim_lstm_layer_f32(_ptr(cur), cur.shape[-1], *[_ptr(t) for t in w], B, T, hidden, _ptr(out), _ptr(ws), nws,
                                          _stream()), "dim_lstm_layer_f32")
        cur = out
    return cur


def split_planes(x, planes):
    """fp32 (rows,K) -> bf16 plane matrix (rows, planes*Kp), Kp = K rounded up to 64 (zero padded)."""
    _req(x, torch.float32, "x")
    rows, K = x.shape
    kp = (K + 63) // 64 * 64
    out = torch.empty(rows, planes * kp, dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.load().dim_split_bf16_planes(_ptr(x), K, rows, K, planes, _ptr(out), _stream()), "dim_split_bf16_planes")
    return out


def linear_tc(a, w, bias=None, residual=None, act=0, slope=0.0, planes=3):
    """Same contract as linear() but on the tcgen05 tensor cores via the bf16 plane split (planes=1: plain bf16)."""
    M, K = a.shape
    N = w.shape[0]
    ap, wp = split_planes(a, planes), split_planes(w, planes)
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().dim_linear_bf16_planes(_ptr(ap), _ptr(wp), K, planes, _ptr(bias), _ptr(residual), N, _ptr(out), N,
                                                  M, N, act, float(slope), _stream()), "dim_linear_bf16_planes")
    return out


def repack_conv_weight(w_oik):
    _req(w_oik, torch.float32, "w")
    co, ci, k = w_oik.shape
    assert k == 5
    out = torch.empty(co, 5, ci, dtype=torch.float32, device=w_oik.device)
    _lib.check(_lib.load().dim_repack_conv_weight(_ptr(w_oik), _ptr(out), co, ci, _stream()), "dim_repack_conv_weight")
    return out


def conv5_leaky(x_btc, w_repacked, bias, slope, lens=None):
    _req(x_btc, torch.float32, "x")
    B, T, Cc = x_btc.shape
    y = torch.empty_like(x_btc)
    _lib.check(_lib.load().dim_conv5_leaky_f32(_ptr(x_btc), _ptr(w_repacked), _ptr(bias), _ptr(lens), _ptr(y), B, T, Cc,
                                               float(slope), _stream()), "dim_conv5_leaky_f32")
    return y


def instance_norm_(x_btc, lens=None, eps=1e-5):
    _req(x_btc, torch.float32, "x")
    B, T, Cc = x_btc.shape
    _lib.check(_lib.load().dim_instance_norm_f32(_ptr(x_btc), _ptr(lens), B, T, Cc, float(eps), _stream()),
               "dim_instance_norm_f32")
    return x_btc


def layer_norm(x, gain, bias=None, eps=1e-5):
    _req(x, torch.float32, "x")
    dim = x.shape[-1]
    rows = x.numel() // dim
    y = torch.empty_like(x)
    _lib.check(_lib.load().dim_layer_norm_f32(_ptr(x), _ptr(gain), _ptr(bias), _ptr(y), rows, dim, float(eps), _stream()),
               "dim_layer_norm_f32")
    return y


def attention(qkv, heads, dim_head, scale, key_mask=None, lens=None, causal=False):
    """qkv (B,T,3*heads*dim_head) laid out '(qkv h d)' -> (B,T,heads*dim_head)."""
    _req(qkv, torch.float32, "qkv")
    B, T, W = qkv.shape
    inner = heads * dim_head
    assert W == 3 * inner
    out = torch.empty(B, T, inner, dtype=torch.float32, device=qkv.device)
    base = qkv.data_ptr()
    _lib.check(_lib.load().dim_attention_f32(base, W, base + 4 * inner, W, base + 8 * inner, W, _ptr(out), inner,
                                             _ptr(key_mask), _ptr(lens), B, heads, T, T, dim_head, float(scale),
                                             int(causal), _stream()), "dim_attention_f32")
    return out


def resample_window_mean(x, factor=0.6):
    """vico_preprocessing.downsample_mean (code/vico_preprocessing.py:7-19): x (t,d) fp32 -> (int(t*factor), d); each output frame
    is the mean of `int(t / new_t)` consecutive input frames starting at i*window (window = 1 at factor 0.6: the reference's
    50 -> 30 fps "downsampling" keeps the first 60 % of the frames)."""
    _req(x, torch.float32, "x")
    t, d = x.shape
    new_t = int(t * factor)
    window = int(t / new_t)
    out = torch.empty(new_t, d, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().dim_resample_features(_ptr(x), t, d, new_t, window, 0, _ptr(out), _stream()), "dim_resample_features")
    return out


def resample_linear(x, new_t):
    """dataset/l2l.downsample_mean (code/dataset/l2l.py:23-29): linear interpolation, align_corners=True.  x (t,d) -> (new_t,d)."""
    _req(x, torch.float32, "x")
    t, d = x.shape
    out = torch.empty(new_t, d, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().dim_resample_features(_ptr(x), t, d, new_t, 1, 1, _ptr(out), _stream()), "dim_resample_features")
    return out


def launch_count() -> int:
    return int(_lib.load().dim_launch_count())
