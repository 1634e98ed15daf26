"""Prompt log-mel front end.

CPU: the oracle restatement (oracle/prompt_mel_oracle.py) against tests/golden/prompt_mel.npz, which
oracle/make_golden_prompt_mel.py produced by calling the UNMODIFIED reference `mel_spectrogram`
(third_party/Matcha-TTS/matcha/utils/audio.py:45-82); the host-side mirror's integer bookkeeping and error behaviour.
GPU: `cv2_prompt_mel` through the reference-shaped host (cosyvoice2_eu_b200/frontend.py) against the same golden vectors,
against the oracle at full size (64 prompts of up to 30 s) and through size-independent properties.

Tolerance: the feature is the log of an fp32 spectrum; the reference's own fp32 FFT differs from the float64 oracle by up to
1.2e-5 on these vectors.  MEL_TOL = 1e-3 max-abs on the log-mel (10x tighter than north_star's 1e-2 for the generated mel)."""
import numpy as np
import pytest
import torch

import prompt_mel_oracle as PO

MEL_TOL = 1e-3
CASES = ["voiced", "ragged", "short", "silence", "loud"]


# ------------------------------------------------------------------------------------------ CPU: oracle pinned
def test_oracle_mel_basis_matches_third_party_filterbank(golden):
    g = golden("prompt_mel")
    mb = PO.mel_basis()
    assert mb.shape == (80, 961) and mb.dtype == np.float32
    assert np.abs(mb - g["mel_basis"]).max() < 1e-8
    # published known answer (librosa.filters.mel docstring, sr=22050, n_fft=2048): second bin of the first band is 0.016
    assert abs(float(PO.mel_basis(22050, 2048, 128, 0.0, 11025.0)[0, 1]) - 0.016) < 5e-4


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(golden, name):
    g = golden("prompt_mel")
    mel = PO.mel_spectrogram(g[name + ".wav"])[0]
    ref = g[name + ".mel"]
    assert mel.shape == ref.shape == (80, PO.n_frames(len(g[name + ".wav"])))
    assert np.abs(mel - ref).max() < 5e-5


def test_frame_count_and_alignment_are_integer_exact():
    from cosyvoice2_eu_b200 import align_prompt
    for L in (721, 959, 960, 10007, 24000 * 30):
        assert PO.n_frames(L) == (L + 1440 - 1920) // 480 + 1
    for feat_T, tok_T in ((65, 40), (65, 20), (64, 32), (3, 9), (1, 5)):
        feat = torch.zeros(1, feat_T, 80)
        tok = torch.zeros(1, tok_T, dtype=torch.int32)
        fl, tl = torch.tensor([feat_T], dtype=torch.int32), torch.tensor([tok_T], dtype=torch.int32)
        f2, fl2, t2, tl2 = align_prompt(feat, fl, tok, tl)
        want = PO.align_prompt(feat_T, tok_T)
        assert (f2.shape[1], t2.shape[1]) == want == (int(fl2[0]), int(tl2[0]))
        assert fl2 is fl and tl2 is tl          # updated in place, like the reference


def test_host_rejects_cpu_tensors_and_other_configs():
    from cosyvoice2_eu_b200 import mel_spectrogram
    from cosyvoice2_eu_b200.lib import Cv2Error
    with pytest.raises(Cv2Error):
        mel_spectrogram(torch.zeros(1, 4800))                       # no CPU fallback
    with pytest.raises(Cv2Error):
        mel_spectrogram(torch.zeros(1, 4800), n_fft=1024)           # only the CosyVoice2 feat_extractor is built


def test_oracle_resample_matches_torchaudio_golden(golden):
    """torchaudio.transforms.Resample(16000, 24000) is what frontend.py:495 calls; the golden entries are its outputs."""
    g = golden("prompt_mel")
    assert np.array_equal(PO.resample_kernel(), g["resample.kernel"])          # the 3 x 16 polyphase filter, bit for bit
    for name in ("rs_a", "rs_odd", "rs_tiny"):
        y = PO.resample_16k_to_24k(g[name + ".in"])[0]
        assert y.shape == g[name + ".out"].shape == (PO.resample_len(len(g[name + ".in"])),)
        assert np.abs(y - g[name + ".out"]).max() < 1e-6


# ------------------------------------------------------------------------------------------ GPU: parity through the C ABI
@pytest.mark.gpu
def test_gpu_resample_matches_torchaudio_golden_and_chain(golden):
    from cosyvoice2_eu_b200 import extract_speech_feat, resample_16k_to_24k
    g = golden("prompt_mel")
    for name in ("rs_a", "rs_odd", "rs_tiny"):
        y = resample_16k_to_24k(torch.from_numpy(g[name + ".in"])[None])
        ref = g[name + ".out"]
        assert tuple(y.shape) == (1, len(ref))
        assert np.abs(y[0].cpu().numpy() - ref).max() < 1e-6
    # the chain a request runs (frontend.py:495-496): 16 kHz prompt -> 24 kHz -> log-mel, all on the device
    y = resample_16k_to_24k(torch.from_numpy(g["rs_a.in"])[None])
    feat, feat_len = extract_speech_feat(y)
    assert np.abs(feat[0].cpu().numpy() - g["rs_a.mel"].T).max() < MEL_TOL
    # ragged batch == single rows, bit for bit; padding is zero
    xs = [g["rs_a.in"], g["rs_odd.in"], g["rs_tiny.in"]]
    host = torch.zeros(3, 16000)
    for i, x in enumerate(xs):
        host[i, :len(x)] = torch.from_numpy(x)
    yb, nb = resample_16k_to_24k(host, lengths=torch.tensor([len(x) for x in xs], dtype=torch.int32))
    assert nb.tolist() == [PO.resample_len(len(x)) for x in xs]
    for i, x in enumerate(xs):
        one = resample_16k_to_24k(torch.from_numpy(x)[None])[0]
        assert torch.equal(yb[i, :len(one)], one) and not yb[i, len(one):].any()

@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_matches_reference_golden(golden, name):
    from cosyvoice2_eu_b200 import extract_speech_feat, mel_spectrogram
    g = golden("prompt_mel")
    wav = torch.from_numpy(g[name + ".wav"])[None]
    feat, feat_len = extract_speech_feat(wav)
    ref = g[name + ".mel"].T                                         # [T, 80]
    assert tuple(feat.shape) == (1,) + ref.shape and int(feat_len[0]) == ref.shape[0]
    err = np.abs(feat[0].cpu().numpy() - ref).max()
    assert err < MEL_TOL, err
    m2 = mel_spectrogram(wav.cuda())                                 # the reference's own layout [B, 80, T]
    assert torch.equal(m2[0].T.contiguous(), feat[0])


@pytest.mark.gpu
def test_gpu_batch_equals_single_prompts_bit_exact(golden):
    from cosyvoice2_eu_b200 import extract_speech_feat, extract_speech_feat_batch
    g = golden("prompt_mel")
    waves = [torch.from_numpy(g[n + ".wav"]) for n in CASES]
    feat, feat_len = extract_speech_feat_batch(waves)
    assert feat_len.tolist() == [PO.n_frames(len(w)) for w in waves]
    for i, w in enumerate(waves):
        one, _ = extract_speech_feat(w[None])
        T = one.shape[1]
        assert torch.equal(feat[i, :T], one[0])
        assert not feat[i, T:].any()                                 # padding rows are zero


@pytest.mark.gpu
def test_gpu_full_size_against_oracle_and_shift_property():
    """64 prompts of 3..30 s (the longest the frontend accepts, frontend.py:466-467 asserts <= 30 s)."""
    from cosyvoice2_eu_b200 import extract_speech_feat_batch
    rng = np.random.Generator(np.random.Philox(key=77))
    lens = [int(24000 * d) for d in rng.uniform(3.0, 30.0, size=63)] + [24000 * 30]
    waves = []
    for L in lens:
        t = np.arange(L) / 24000.0
        f0 = rng.uniform(90, 250)
        y = sum(np.sin(2 * np.pi * h * f0 * t + rng.uniform(0, 6.28)) / h for h in range(1, 24))
        y = 0.5 * y / np.abs(y).max() * (0.55 + 0.45 * np.sin(2 * np.pi * 2.3 * t)) + 0.01 * rng.standard_normal(L)
        waves.append(torch.from_numpy(y.astype(np.float32)))
    feat, feat_len = extract_speech_feat_batch(waves)
    assert feat_len.tolist() == [PO.n_frames(L) for L in lens]
    for i in (0, 17, 63):                                            # the oracle on three of them (seconds of CPU)
        ref = PO.mel_spectrogram(waves[i].numpy())[0].T
        err = np.abs(feat[i, :ref.shape[0]].cpu().numpy() - ref).max()
        assert err < MEL_TOL, (i, err)
    # shifting every prompt by exactly one hop shifts the interior frames by one, bit for bit (same samples, same order)
    shifted = [w[480:] for w in waves]
    feat_s, _ = extract_speech_feat_batch(shifted)
    for i in (0, 17, 63):
        T = PO.n_frames(lens[i] - 480)
        assert torch.equal(feat_s[i, 2:T - 2], feat[i, 3:T - 1])
