"""Error behaviour of the C ABI on a GPU box (include/dimb200.h: integer status + dim_last_error(), no exceptions across the
boundary, nothing computed on bad input): missing weights, wrong shapes, short workspaces, bad arguments."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200 import _lib  # noqa: E402
from dim_b200.schema import S2SConfig, VQConfig  # noqa: E402

DIM_EINVAL, DIM_EMISSING, DIM_EWORKSPACE = 1, 4, 5


def _err(lib):
    return (lib.dim_last_error() or b"").decode()


def test_missing_and_misshapen_weights_are_reported(vq_sd):
    from dim_b200.engine import Handle, VQEngine
    h = Handle()
    sd = dict(vq_sd)
    sd.pop("quantize.embedding.weight")
    h.register(sd)
    with pytest.raises(RuntimeError, match="quantize.embedding.weight"):
        VQEngine(h, VQConfig())
    h2 = Handle()
    sd = dict(vq_sd)
    sd["encoder.vertice_mapping.0.weight"] = torch.zeros(384, 60)
    h2.register(sd)
    with pytest.raises(RuntimeError, match="shape mismatch"):
        VQEngine(h2, VQConfig())


def test_short_workspace_and_bad_model_are_refused(vq_sd):
    from dim_b200.engine import Handle, VQEngine
    h = Handle()
    h.register(vq_sd)
    eng = VQEngine(h, VQConfig())
    lib = h.lib
    B, T = 2, 16
    x = torch.randn(B, T, 56, device="cuda")
    idx = torch.full((B, T), -7, dtype=torch.int64, device="cuda")
    need = lib.dim_vqvae_workspace_bytes(h.h, eng.model, B, T)
    ws = torch.empty(need // 2, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    rc = lib.dim_vqvae_encode(h.h, eng.model, x.data_ptr(), None, None, B, T, idx.data_ptr(), None, None, ws.data_ptr(), need // 2, s)
    assert rc == DIM_EWORKSPACE and "workspace" in _err(lib)
    torch.cuda.synchronize()
    assert int((idx != -7).sum()) == 0                                  # nothing was launched
    rc = lib.dim_vqvae_encode(h.h, 99, x.data_ptr(), None, None, B, T, idx.data_ptr(), None, None, ws.data_ptr(), need // 2, s)
    assert rc == DIM_EINVAL and "bad model" in _err(lib)
    assert lib.dim_vqvae_workspace_bytes(h.h, 99, B, T) == 0
    rc = lib.dim_vqvae_decode(h.h, eng.model, None, None, None, B, T, x.data_ptr(), ws.data_ptr(), need // 2, s)
    assert rc == DIM_EINVAL                                             # neither codes nor quant


def test_operator_argument_checks():
    lib = _lib.load()
    s = torch.cuda.current_stream().cuda_stream
    z = torch.randn(8, 100, device="cuda")
    E = torch.randn(512, 100, device="cuda")
    o = torch.empty(8, dtype=torch.int64, device="cuda")
    assert lib.dim_vq_argmin(z.data_ptr(), E.data_ptr(), o.data_ptr(), 8, 100, 512, s) == DIM_EINVAL      # D must be 64/128/256
    assert "D must be" in _err(lib)
    x = torch.randn(10, 6, device="cuda")
    y = torch.empty(6, 6, device="cuda")
    assert lib.dim_resample_features(x.data_ptr(), 10, 6, 6, 1, 0, y.data_ptr(), s) == DIM_EINVAL          # d % 4 != 0
    x = torch.randn(10, 8, device="cuda")
    y = torch.empty(6, 8, device="cuda")
    assert lib.dim_resample_features(x.data_ptr(), 10, 8, 6, 2, 0, y.data_ptr(), s) == DIM_EINVAL          # 6 windows of 2 > 10 frames
    assert lib.dim_resample_features(x.data_ptr(), 10, 8, 6, 1, 0, y.data_ptr(), s) == 0


def test_generate_argument_checks(slmft_sd):
    from dim_b200.engine import PREC_FP32_TC, Handle, SLMFTEngine
    h = Handle()
    h.register(slmft_sd)
    s2s = SLMFTEngine(h, S2SConfig(), precision=PREC_FP32_TC)
    B, T = 2, 8
    c = dim_b200.synth.make_clips(B, T, seed=1)
    ctx = s2s.context(c["v_speaker"].cuda(), c["v_audio"].cuda(), c["mask"].cuda())
    prompt = torch.zeros(B, dtype=torch.int64, device="cuda")
    with pytest.raises(RuntimeError, match="uniforms"):                  # sampling without uniforms
        s2s.generate(ctx, c["mask"].cuda(), prompt, T - 1, temperature=1.0, uniforms=None)
    lib = h.lib
    out = torch.empty(B, T - 1, dtype=torch.int64, device="cuda")
    m8 = c["mask"].to(torch.uint8).cuda()
    s = torch.cuda.current_stream().cuda_stream
    rc = lib.dim_slmft_generate(h.h, s2s.model, ctx.data_ptr(), m8.data_ptr(), prompt.data_ptr(), B, T, T - 1, C.c_float(-1.0), 52,
                                None, out.data_ptr(), None, None, 0, s)
    assert rc == DIM_EINVAL and "temperature" in _err(lib)
    rc = lib.dim_slmft_generate(h.h, s2s.model, ctx.data_ptr(), m8.data_ptr(), prompt.data_ptr(), B, T, T - 1, C.c_float(0.0), 0,
                                None, out.data_ptr(), None, None, 0, s)
    assert rc == DIM_EWORKSPACE
