"""VQ-VAE encode -> quantize -> decode on the GPU vs the oracle and vs the golden vectors minted from the real reference.

Bar (BASELINE.json north_star): bit-exact code indices; decoded FLAME coefficients within 1e-4 abs (fp32)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from dim_b200.schema import VQConfig  # noqa: E402
from oracle import vqvae as OV  # noqa: E402

CFG = VQConfig()
TOL = 1e-4


@pytest.fixture(scope="module", params=["fp32_ffma", "fp32_tcgen05"])
def engine(vq_sd, request):
    """Both fp32-grade arithmetic modes must meet the same bar: FFMA kernels, and tcgen05 with the exact 3-plane bf16 split."""
    from dim_b200.engine import PREC_FP32, PREC_FP32_TC, Handle, VQEngine
    h = Handle()
    h.register(vq_sd)
    return VQEngine(h, CFG, precision=PREC_FP32 if request.param == "fp32_ffma" else PREC_FP32_TC)


def _x(case):
    g = torch.Generator().manual_seed(case["x_seed"])
    return torch.randn(case["B"], case["T"], 56, generator=g) * case["x_scale"]


@pytest.mark.parametrize("name", ["c1_T300_B1", "f4_T64_B3", "short_T5_B2"])
def test_golden_roundtrip(engine, golden, name):
    """BASELINE configs[0]: same seeded clip, indices identical to the REAL reference's, decode within 1e-4."""
    case = golden["cases"][name]
    x = _x(case).cuda()
    idx, z, quant = engine.encode(x, want_z=True, want_quant=True)
    assert torch.allclose(z.cpu(), case["z"], atol=TOL), float((z.cpu() - case["z"]).abs().max())
    assert torch.equal(idx.cpu(), case["idx"]), f"{int((idx.cpu() != case['idx']).sum())} code mismatches"
    dec = engine.decode(quant=quant)
    assert torch.allclose(dec.cpu(), case["dec_idx"], atol=TOL), float((dec.cpu() - case["dec_idx"]).abs().max())
    dec2 = engine.decode(codes=idx)
    assert torch.equal(dec2, dec)          # fused gather == explicit (B,C,L) quant input, bit for bit


def test_quant_is_exact_codebook_rows(engine, vq_sd):
    x = (torch.randn(2, 33, 56, generator=torch.Generator().manual_seed(3)) * 0.3).cuda()
    idx, _, quant = engine.encode(x, want_quant=True)
    rows = vq_sd["quantize.embedding.weight"][idx.cpu()]             # (B,T,D)
    assert torch.equal(quant.cpu(), rows.permute(0, 2, 1))


def test_batch_index_quirk(engine, vq_sd):
    """sample b gets pe[b]; a shard must be able to pass global batch positions."""
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 48, 56, generator=g) * 0.3
    full_idx, full_z, _ = engine.encode(x.cuda(), want_z=True)
    ref = OV.encode(vq_sd, x, CFG)[2][2].view(4, 48)
    assert torch.equal(full_idx.cpu(), ref)
    # shard [2:4] with explicit global indices == rows 2:4 of the full batch
    bi = torch.tensor([2, 3], dtype=torch.int32).cuda()
    part_idx, part_z, _ = engine.encode(x[2:4].cuda(), batch_index=bi, want_z=True)
    assert torch.equal(part_idx, full_idx[2:4]) and torch.equal(part_z, full_z[2:4])
    # and without it the result differs (the quirk is real)
    naive_idx, naive_z, _ = engine.encode(x[2:4].cuda(), want_z=True)
    assert not torch.equal(naive_z, full_z[2:4])
    dec_full = engine.decode(codes=full_idx)
    dec_part = engine.decode(codes=full_idx[2:4].contiguous(), batch_index=bi)
    assert torch.equal(dec_part, dec_full[2:4])
    ref_dec = OV.decode_indices(vq_sd, ref, CFG)
    assert torch.allclose(dec_full.cpu(), ref_dec, atol=TOL)


def test_ragged_lengths_match_per_sample_encode(engine, vq_sd):
    """SLMFT.forward_vq encodes each sample's valid prefix alone with B=1 (seq2seq_pretrain.py:484-491)."""
    g = torch.Generator().manual_seed(9)
    B, T = 3, 50
    lens = torch.tensor([50, 31, 5], dtype=torch.int32)
    x = torch.randn(B, T, 56, generator=g) * 0.3
    zero = torch.zeros(B, dtype=torch.int32).cuda()                  # every sample is "batch slot 0"
    idx, z, _ = engine.encode(x.cuda(), lens=lens.cuda(), batch_index=zero, want_z=True)
    for i in range(B):
        L = int(lens[i])
        zr = OV.encoder(vq_sd, x[i:i + 1, :L], CFG)
        ir = OV.encode(vq_sd, x[i:i + 1, :L], CFG)[2][2].view(-1)
        assert torch.allclose(z[i, :L].cpu(), zr[0], atol=TOL), float((z[i, :L].cpu() - zr[0]).abs().max())
        assert torch.equal(idx[i, :L].cpu(), ir)


def test_full_size_properties(engine, vq_sd):
    """BASELINE configs[2] size (B=256, T=300): size-independent properties instead of a CPU run.
       (1) batch slot 0 equals the single-clip result, (2) decode(encode) is deterministic run to run,
       (3) every code is in range and decode of codes equals decode of the gathered rows."""
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(256, 300, 56, generator=g) * 0.3).cuda()
    idx, _, quant = engine.encode(x, want_quant=True)
    idx1, _, _ = engine.encode(x[:1].contiguous())
    assert torch.equal(idx[:1], idx1)
    assert int(idx.min()) >= 0 and int(idx.max()) < 512
    dec = engine.decode(codes=idx)
    assert torch.equal(dec, engine.decode(quant=quant))
    assert torch.equal(dec, engine.decode(codes=idx))
    ref0 = OV.decode_indices(vq_sd, idx[:1].cpu(), CFG)
    assert torch.allclose(dec[:1].cpu(), ref0, atol=TOL)
    assert torch.isfinite(dec).all()


def test_bf16_decoder_is_close(vq_sd):
    """bf16 mode (BASELINE configs[2] 'bf16 fused transformer + VQ decode'): the codes -> frames decoder with plain bf16 GEMM
    operands is NOT a parity mode; it must stay within bf16 rounding noise of the oracle (outputs are O(1))."""
    from dim_b200.engine import PREC_BF16, Handle, VQEngine
    h = Handle()
    h.register(vq_sd)
    eng = VQEngine(h, CFG, precision=PREC_BF16)
    codes = torch.randint(0, 512, (3, 40), generator=torch.Generator().manual_seed(5))
    dec = eng.decode(codes=codes.cuda())
    ref = OV.decode_indices(vq_sd, codes, CFG)
    err = float((dec.cpu() - ref).abs().max())
    assert torch.isfinite(dec).all() and err < 5e-2 * max(1.0, float(ref.abs().max())), err
