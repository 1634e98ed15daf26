"""x-vector features (second half): kaldi.fbank(num_mel_bins=80, dither=0, 16 kHz) minus its column mean
(cosyvoice/cli/frontend.py:276-278).  tests/golden/fbank.npz holds torchaudio.compliance.kaldi.fbank's own outputs on seeded
prompts (oracle/make_golden_fbank.py); the numpy oracle and the CUDA kernel are both checked against them.

Tolerance: torchaudio works in fp32 through an FFT, so bins whose power sits within a few decades of the fp32 epsilon floor
(log value < -10: digital silence, the stop band of a clean tone) carry its own rounding noise of ~1e-3; everywhere else the
float64 oracle agrees with it to 1e-4 and the fp32 kernel (two fp32 computations, each ~6e-5 from the float64 truth) to 3e-4
on log-mel values that span +-10."""
import numpy as np
import pytest
import torch

import fbank_oracle as F

CASES = ["glide", "noise", "ragged", "one_frame", "silence", "clipped"]


def _check(got, g, name, tol_hi=1e-4, tol_lo=5e-3):
    raw, want = g[f"{name}.fbank_raw"], g[f"{name}.feat"]
    assert got.shape == want.shape
    d = np.abs(got - want)
    strong = raw.min(axis=0, keepdims=True) > -10.0           # columns that never come near the floor: the mean is clean too
    strong = np.broadcast_to(strong, raw.shape)
    if strong.any():
        assert d[strong].max() <= tol_hi, (name, d[strong].max())
    assert d.max() <= tol_lo, (name, d.max())


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_torchaudio_kaldi_fbank(golden, name):
    g = golden("fbank")
    _check(F.fbank(g[f"{name}.wav"]), g, name)
    assert F.num_frames(g[f"{name}.wav"].size) == g[f"{name}.feat"].shape[0]


def test_mel_banks_match_torchaudio():
    kaldi = pytest.importorskip("torchaudio.compliance.kaldi")
    ref, _ = kaldi.get_mel_banks(80, 512, 16000.0, 20.0, 0.0, 100.0, -500.0, 1.0)
    got = F.mel_banks()
    assert got.shape == (80, 257) and np.all(got[:, 256] == 0)
    assert np.abs(got[:, :256] - ref.numpy()).max() < 5e-5       # torchaudio builds the slopes in fp32, the oracle in fp64


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_fbank_matches_torchaudio(golden, name):
    from cosyvoice2_eu_b200 import extract_spk_feat
    g = golden("fbank")
    feat = extract_spk_feat(torch.from_numpy(g[f"{name}.wav"])[None])
    _check(feat[0].cpu().numpy(), g, name, tol_hi=3e-4)


@pytest.mark.gpu
def test_gpu_fbank_ragged_batch_equals_single_prompts(golden):
    from cosyvoice2_eu_b200 import extract_spk_feat
    g = golden("fbank")
    waves = [torch.from_numpy(g[f"{n}.wav"]) for n in CASES]
    lens = torch.tensor([w.numel() for w in waves], dtype=torch.int32)
    batch = torch.zeros(len(waves), int(lens.max()))
    for i, w in enumerate(waves):
        batch[i, :w.numel()] = w
    feat, feat_len = extract_spk_feat(batch, lengths=lens)
    assert feat_len.tolist() == [F.num_frames(int(n)) for n in lens]
    for i, n in enumerate(CASES):
        one = extract_spk_feat(waves[i][None])[0]
        k = int(feat_len[i])
        assert torch.equal(feat[i, :k], one)
        assert float(feat[i, k:].abs().max()) == 0.0 if k < feat.shape[1] else True
        _check(feat[i, :k].cpu().numpy(), g, n, tol_hi=3e-4)


@pytest.mark.gpu
def test_gpu_fbank_rejects_too_short():
    from cosyvoice2_eu_b200 import extract_spk_feat
    from cosyvoice2_eu_b200.lib import Cv2Error
    with pytest.raises(Cv2Error):
        extract_spk_feat(torch.zeros(1, 399))
