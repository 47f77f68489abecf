"""GPU-side loaders (dim_b200.loader) against goldens minted by the REAL reference loaders + collate functions
(tests/golden/make_loader_golden.py: dataset/data_loader.py ViCoDataset / LmListenerDataset, dataset/l2l.py LmListenerDataset) on the
same seeded fixtures: shapes, lengths, names, sums and a strided sample of every padded batch tensor."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

import dim_b200  # noqa: E402
from dim_b200 import l2l_artifacts as A  # noqa: E402
from dim_b200 import loader  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(ROOT, "tests", "golden", "loader_reference.pt"), weights_only=False)


@pytest.fixture(scope="module")
def fixtures(gold, tmp_path_factory):
    root = tmp_path_factory.mktemp("data_root")
    f = gold["fixtures"]
    clips = dim_b200.synth.make_clips(f["vico"]["batch"], f["vico"]["frames"], seed=f["vico"]["seed"], ragged=True)
    clips = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in clips.items()}
    A.write_vico_fixtures(str(root / "data"), clips, split=f["vico"]["split"])
    A.write_lm_listener_fixtures(str(root / "data" / "lm"), mode="test", seed=f["lm"]["seed"], lengths=tuple(f["lm"]["lengths"]))
    return root


def _check(t, ref, exact=True):
    assert tuple(t.shape) == ref["shape"]
    f = t.cpu().reshape(-1)
    if exact:
        assert torch.equal(f[::97], ref["sample"])
        assert float(f.double().sum()) == ref["sum"]
    else:
        assert float((f[::97] - ref["sample"]).abs().max()) < 2e-6
        assert abs(float(f.double().sum()) - ref["sum"]) < 1e-3 * max(1.0, abs(ref["sum"]))


def test_vico_batches_equal_the_reference_loader(gold, fixtures):
    ds = loader.ViCoClips(str(fixtures / "data" / "vico_processed_30fps"), str(fixtures / "data" / "RLD_data.csv"), mode="test")
    got = list(ds.batches(3))
    assert len(got) == len(gold["vico"])
    for (src, tgt, lens, ids, names, mask), ref in zip(got, gold["vico"]):
        assert src.is_cuda and lens == ref["lens"] and [os.path.basename(n) for n in names] == ref["names"]
        assert ids[0].tolist() == ref["speaker_ids"] and ids[1].tolist() == ref["listener_ids"]
        _check(src, ref["src"])
        _check(tgt, ref["tgt"])
        exp_mask = torch.arange(src.shape[1])[None, :] < torch.tensor(lens)[:, None]                 # x_engine_pt.py:246-249
        assert torch.equal(mask.cpu(), exp_mask)
        assert bool((src[..., :56][mask] == 1).all()) and bool((src[~mask] == 0).all())


@pytest.mark.parametrize("hubert", [False, True])
def test_lm_listener_batches_equal_the_reference_loader(gold, fixtures, hubert):
    ds = loader.LmListenerSegments(str(fixtures / "data" / "lm"), mode="test", hubert=hubert)
    ref_b = gold["lm_hubert" if hubert else "lm_zeros"]
    got = list(ds.batches(2))
    assert len(got) == len(ref_b) and len(ds) == 3                      # 30, [10 dropped], 1100 -> one 1024-frame chunk, 64
    for (src, tgt, xl, yl, names, mask), ref in zip(got, ref_b):
        assert xl == ref["lens"] and names == ref["names"]
        _check(tgt, ref["tgt"])
        _check(src, ref["src"], exact=not hubert)                       # the device-side linear resampling: within 1e-6
        if not hubert:
            assert bool((src[..., 56:] == 0).all())


def test_assemble_batch_edge_cases():
    g = torch.Generator().manual_seed(0)
    lens = [5, 1, 9]
    li = [torch.randn(n, 56, generator=g).numpy() for n in lens]
    au = [torch.randn(n, 768, generator=g).numpy() for n in lens]
    src, tgt, mask = loader.assemble_batch(None, au, li, lens, T=12)
    assert src.shape == (3, 12, 824) and mask.sum().item() == sum(lens)
    for b, n in enumerate(lens):
        assert torch.equal(tgt[b, :n].cpu(), torch.from_numpy(li[b])) and bool((tgt[b, n:] == 0).all())
        assert torch.equal(src[b, :n, 56:].cpu(), torch.from_numpy(au[b])) and bool((src[b, n:] == 0).all())
