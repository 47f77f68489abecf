"""Generate tests/golden/vq_speaker_reference.pt by running the REAL reference VQSpeakerAutoEncoder
(/root/reference/code/models/stage1_BIWI.py:140-251, built by models.get_model from code/config_speaker_old.yaml:
arch stage1_BIWI_speaker, in_dim 824, hidden 768, 8 codes per frame).

Run in the build container only (/root/reference does not exist on the GPU box):
    python tests/golden/make_speaker_golden.py

The synthetic state_dict (dim_b200.synth.make_vqspeaker_state_dict, seed stored) is loaded with strict=True; the script asserts that
oracle/vqvae.py agrees with the reference (indices identical, floats <= 1e-5) and stores the reference outputs.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/code"

import dim_b200  # noqa: E402
from dim_b200.schema import SPEAKER_VQ, VQConfig  # noqa: E402
from oracle import vqvae as O  # noqa: E402
from make_golden import sd_checksum  # noqa: E402


def main():
    sys.path.insert(0, REF)
    from base import config
    from models import get_model
    rcfg = config.load_cfg_from_cfg_file(os.path.join(REF, "config_speaker_old.yaml"))
    assert rcfg.arch == "stage1_BIWI_speaker"
    cfg = VQConfig.from_cfg(rcfg)
    assert cfg == SPEAKER_VQ, cfg
    seed_w = 137
    sd = dim_b200.synth.make_vqspeaker_state_dict(seed_w)
    ref = get_model(rcfg).eval()
    assert type(ref).__name__ == "VQSpeakerAutoEncoder"
    ref.load_state_dict(sd, strict=True)
    out = {"weights_seed": seed_w, "weights_sha256": sd_checksum(sd), "torch": torch.__version__, "cases": {}}
    for name, B, T, seed_x in (("sp_T40_B2", 2, 40, 21), ("sp_T24_B3", 3, 24, 22)):
        g = torch.Generator().manual_seed(seed_x)
        x = torch.randn(B, T, cfg.in_dim, generator=g) * 0.3
        with torch.no_grad():
            quant, loss, (ppl, onehot, idx) = ref.encode(x)
            dec = ref.decode(quant)
            z = ref.encoder(x)
            dec_idx = ref.decode_to_img(idx, (B, T * cfg.face_quan_num, cfg.zquant_dim))
            q2, l2, (p2, oh2, i2) = O.encode(sd, x, cfg)
            assert torch.equal(i2, idx), name
            pairs = [(q2, quant), (l2, loss), (O.speaker_decode(sd, quant, cfg), dec), (O.encoder(sd, x, cfg), z)]
            err = max(float((a - b).abs().max()) for a, b in pairs)
            assert err <= 1e-5, (name, err)
            d = O.distances(z.reshape(-1, cfg.zquant_dim), sd["quantize.embedding.weight"])
            top2 = torch.topk(d, 2, dim=1, largest=False).values
        assert quant.shape == (B, cfg.zquant_dim, T * cfg.face_quan_num) and dec.shape == (B, T, 824)
        out["cases"][name] = dict(B=B, T=T, x_seed=seed_x, x_scale=0.3, idx=idx.view(B, T * cfg.face_quan_num).clone(),
                                  z=z.clone(), dec=dec.clone(), dec_idx=dec_idx.clone(), loss=loss.clone(),
                                  top2_gap_min=float((top2[:, 1] - top2[:, 0]).min()), oracle_max_err=err)
        print(name, "codes", len(idx.unique()), "min top-2 gap", out["cases"][name]["top2_gap_min"], "oracle err", err)
    path = os.path.join(HERE, "vq_speaker_reference.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
