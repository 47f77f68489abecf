"""Generate tests/golden/emoca_reference.pt by running the REAL reference EmocaConverter class
(/root/reference/code/seq2seq_pretrain.py:759-832: frozen 56-d speaker VQ-VAE -> nn.LSTM(56, 384, 2 layers, bidirectional) ->
Linear(768,768) + LeakyReLU -> Linear(768,70110) + template) on the CPU.

Run in the build container only (/root/reference does not exist on the GPU box):
    python tests/golden/make_emoca_golden.py

The class reads ./config.yaml and ./runs_speaker_new/_RANK0/model/model.pth.tar from the working directory (:762-772): the script
builds such a directory under /tmp from the reference's config.yaml and the synthetic speaker VQ checkpoint.  seq2seq_pretrain.py
imports x_transformers at module level (absent offline, and not used by EmocaConverter): a stub module with the imported names
stands in for the import only.  The synthetic state_dict (dim_b200.synth.make_emoca_converter_state_dict, seed stored) is loaded with
strict=True; the script asserts that oracle/speaker_mesh.py agrees with the reference and stores the reference outputs (the
(B,T,70110) output strided by 997 plus its sum; weights are regenerated from the seed, a checksum detects generator drift).
"""
import os
import shutil
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
REF = "/root/reference/code"

import dim_b200  # noqa: E402
from dim_b200.schema import VQConfig  # noqa: E402
from oracle import speaker_mesh as OM  # noqa: E402
from make_golden import sd_checksum  # noqa: E402


def main():
    stub = types.ModuleType("x_transformers")
    for n in ("TransformerWrapper", "ContinuousTransformerWrapper", "Encoder", "Decoder", "AutoregressiveWrapper",
              "ContinuousAutoregressiveWrapper"):
        setattr(stub, n, type(n, (), {}))
    sys.modules["x_transformers"] = stub
    sys.path.insert(0, REF)
    import seq2seq_pretrain as R

    seed_w, seed_x, stride = 131, 17, 997
    sd = dim_b200.synth.make_emoca_converter_state_dict(seed_w)
    work = tempfile.mkdtemp(prefix="emoca_golden_")
    shutil.copy(os.path.join(REF, "config.yaml"), os.path.join(work, "config.yaml"))
    os.makedirs(os.path.join(work, "runs_speaker_new", "_RANK0", "model"))
    vq = {k[len("speaker_vq."):]: v for k, v in sd.items() if k.startswith("speaker_vq.")}
    torch.save({"state_dict": vq}, os.path.join(work, "runs_speaker_new", "_RANK0", "model", "model.pth.tar"))
    cwd = os.getcwd()
    os.chdir(work)
    try:
        ref = R.EmocaConverter().eval()
    finally:
        os.chdir(cwd)
        shutil.rmtree(work)
    assert set(ref.state_dict().keys()) == set(sd.keys()), set(ref.state_dict().keys()) ^ set(sd.keys())
    ref.load_state_dict(sd, strict=True)

    g = torch.Generator().manual_seed(seed_x)
    B, T = 2, 12
    v = torch.randn(B, T, 56, generator=g) * 0.3
    tpl = torch.randn(B, 70110, generator=g) * 0.1
    with torch.no_grad():
        out, none = ref(None, tpl, v)
        want, dec = OM.emoca_converter_forward(sd, tpl, v, VQConfig())
        err = float((out - want).abs().max())
        assert none is None and err <= 1e-5, err
        # the input path (commented out in EmocaConverter.forward :817-819, live in SpeakerSLMFT.forward :710-713)
        x = ref.squasher(ref.vertice_mapping(out - tpl.unsqueeze(1)).permute(0, 2, 1)).permute(0, 2, 1)
        err2 = float((x - OM.mesh_to_motion(sd, out, tpl)).abs().max())
        assert err2 <= 1e-5, err2
    gold = dict(weights_seed=seed_w, weights_sha256=sd_checksum(sd), x_seed=seed_x, B=B, T=T, stride=stride, torch=torch.__version__,
                out_strided=out[..., ::stride].clone(), out_sum=float(out.double().sum()), out_abs_sum=float(out.double().abs().sum()),
                dec=dec.clone(), motion=x.clone(), oracle_max_err=max(err, err2))
    path = os.path.join(HERE, "emoca_reference.pt")
    torch.save(gold, path)
    print("oracle vs reference max err", err, err2, "-> wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
