"""TEST INFRASTRUCTURE ONLY -- numpy float64 restatement of the x-vector front end of the reference:

    feat = kaldi.fbank(speech, num_mel_bins=80, dither=0, sample_frequency=16000)       cosyvoice/cli/frontend.py:277
    feat = feat - feat.mean(dim=0, keepdim=True)                                        cosyvoice/cli/frontend.py:278

`kaldi.fbank` is torchaudio.compliance.kaldi.fbank (a dependency, torchaudio==2.3.1 in cosy_repo/requirements.txt:36; the image has
2.11 whose kaldi.py is the same algorithm), not code under /root/reference.  Restated from its published algorithm with the
defaults the call leaves in place: 25 ms frames / 10 ms shift (400 / 160 samples), snip_edges, per-frame DC removal, pre-emphasis
0.97 with a replicated first sample, povey window (hann(periodic=False) ** 0.85), zero padding to 512, power spectrum, 80
triangular filters in the kaldi mel scale 1127 ln(1 + f / 700) between 20 Hz and Nyquist (the Nyquist bin gets weight 0), log with
the fp32 epsilon floor.  Pinned against torchaudio itself: tests/golden/fbank.npz (oracle/make_golden_fbank.py).
Only tests/ and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

SR, WIN, SHIFT, NFFT, NMEL = 16000, 400, 160, 512, 80
EPS = float(np.finfo(np.float32).eps)


def num_frames(n_samples):
    return 0 if n_samples < WIN else 1 + (n_samples - WIN) // SHIFT


def povey_window():
    n = np.arange(WIN, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2 * np.pi * n / (WIN - 1))) ** 0.85


def mel_banks():
    """[80, 257] (last column zero), torchaudio.compliance.kaldi.get_mel_banks with vtln_warp = 1."""
    mel = lambda f: 1127.0 * np.log(1.0 + f / 700.0)
    lo, hi = mel(20.0), mel(SR / 2)
    delta = (hi - lo) / (NMEL + 1)
    b = np.arange(NMEL, dtype=np.float64)[:, None]
    left, center, right = lo + b * delta, lo + (b + 1) * delta, lo + (b + 2) * delta
    m = mel(SR / NFFT * np.arange(NFFT // 2, dtype=np.float64))[None, :]
    w = np.maximum(0.0, np.minimum((m - left) / (center - left), (right - m) / (right - center)))
    return np.pad(w, ((0, 0), (0, 1)))


def fbank(wave, subtract_mean=True):
    """wave: 1-D float array at 16 kHz -> [frames, 80] log-mel (float64), column means removed (frontend.py:278)."""
    x = np.asarray(wave, np.float64).reshape(-1)
    m = num_frames(x.size)
    if m == 0:
        return np.zeros((0, NMEL))
    idx = np.arange(m)[:, None] * SHIFT + np.arange(WIN)[None, :]
    fr = x[idx]
    fr = fr - fr.mean(axis=1, keepdims=True)
    prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)
    fr = (fr - 0.97 * prev) * povey_window()[None, :]
    spec = np.abs(np.fft.rfft(fr, n=NFFT, axis=1)) ** 2
    out = np.log(np.maximum(spec @ mel_banks().T, EPS))
    if subtract_mean:
        out = out - out.mean(axis=0, keepdims=True)
    return out
