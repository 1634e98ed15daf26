"""Generate tests/golden/prompt_mel.npz from the UNMODIFIED reference `mel_spectrogram` (container only).

    python oracle/make_golden_prompt_mel.py            # needs /root/reference

MT/utils/audio.py is loaded from where it lies (read-only) and called with the cosyvoice2.yaml:152-160 arguments.  Its one
missing import, `librosa.filters.mel`, is served by `transformers.audio_utils.mel_filter_bank(norm="slaney", mel_scale="slaney")`
-- an independent implementation of the same published filterbank that ships in this image -- so that the stored outputs do
not depend on oracle/prompt_mel_oracle.py's own restatement of it.  The filterbank is stored too (`mel_basis`).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "prompt_mel.npz")
AUDIO_PY = "/root/reference/cosy_repo/third_party/Matcha-TTS/matcha/utils/audio.py"


from transformers.audio_utils import mel_filter_bank  # noqa: E402  (before the librosa stub below exists)


def third_party_mel(sr, n_fft, n_mels, fmin, fmax):
    fb = mel_filter_bank(num_frequency_bins=n_fft // 2 + 1, num_mel_filters=n_mels, min_frequency=float(fmin),
                         max_frequency=float(fmax), sampling_rate=sr, norm="slaney", mel_scale="slaney")
    return np.ascontiguousarray(np.asarray(fb, dtype=np.float64).T.astype(np.float32))     # librosa layout [n_mels, bins]


def load_reference_audio():
    lib = types.ModuleType("librosa")
    filt = types.ModuleType("librosa.filters")
    filt.mel = lambda sr, n_fft, n_mels, fmin, fmax: third_party_mel(sr, n_fft, n_mels, fmin, fmax)
    lib.filters = filt
    sys.modules.setdefault("librosa", lib)
    sys.modules.setdefault("librosa.filters", filt)
    spec = importlib.util.spec_from_file_location("ref_matcha_audio", AUDIO_PY)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_waves():
    """Seeded synthetic prompts: a voiced glide with harmonics and breath noise, a ragged-length one, the shortest legal
    one (721 samples: reflect padding needs > 720), digital silence (the clamp path) and a full-scale square-ish wave."""
    rng = np.random.Generator(np.random.Philox(key=2024))
    def voiced(L, f_lo, f_hi, amp):
        t = np.arange(L) / 24000.0
        f0 = f_lo + (f_hi - f_lo) * 0.5 * (1 - np.cos(2 * np.pi * t / (L / 24000.0)))
        ph = 2 * np.pi * np.cumsum(f0) / 24000.0
        y = np.zeros(L)
        for h in range(1, 40):
            fh = h * f0.mean()
            env = 1.0 / h * (1.0 + 2.0 * np.exp(-((fh - 700.0) / 300.0) ** 2) + 1.5 * np.exp(-((fh - 2300.0) / 500.0) ** 2))
            y += env * np.sin(h * ph + rng.uniform(0, 2 * np.pi))
        y = y / np.abs(y).max() * amp * (0.6 + 0.4 * np.sin(2 * np.pi * 3.1 * t))
        return (y + 0.003 * rng.standard_normal(L)).astype(np.float32)
    return {
        "voiced": voiced(31200, 110.0, 190.0, 0.8),
        "ragged": voiced(10007, 200.0, 260.0, 0.5),
        "short": voiced(721, 150.0, 150.0, 0.3),
        "silence": np.zeros(2400, dtype=np.float32),
        "loud": np.clip(voiced(4800, 90.0, 95.0, 3.0), -1.0, 1.0).astype(np.float32),
    }


def main():
    audio = load_reference_audio()
    out = {"mel_basis": third_party_mel(24000, 1920, 80, 0, 8000)}
    for name, w in make_waves().items():
        with torch.no_grad():
            mel = audio.mel_spectrogram(torch.from_numpy(w)[None], n_fft=1920, num_mels=80, sampling_rate=24000, hop_size=480,
                                        win_size=1920, fmin=0, fmax=8000, center=False)
        out[f"{name}.wav"] = w
        out[f"{name}.mel"] = mel[0].numpy().astype(np.float32)          # [80, T]
        print(name, w.shape, tuple(mel.shape), float(mel.min()), float(mel.max()))
    # 16 kHz -> 24 kHz exactly as CV/cli/frontend.py:495 does it, then the reference mel of the result (the chain a request runs)
    import torchaudio
    rs = torchaudio.transforms.Resample(orig_freq=16000, new_freq=24000)
    out["resample.kernel"] = rs.kernel[:, 0, :].numpy().astype(np.float32)          # [3, 16]
    rng = np.random.Generator(np.random.Philox(key=16000))
    for name, L in (("rs_a", 16000), ("rs_odd", 7777), ("rs_tiny", 5)):
        t = np.arange(L) / 16000.0
        x = (0.5 * np.sin(2 * np.pi * 220.0 * t) + 0.2 * np.sin(2 * np.pi * 3100.0 * t + 1.0) + 0.05 * rng.standard_normal(L)).astype(np.float32)
        with torch.no_grad():
            y = rs(torch.from_numpy(x)[None])
        out[f"{name}.in"] = x
        out[f"{name}.out"] = y[0].numpy().astype(np.float32)
        print(name, x.shape, tuple(y.shape))
        if L >= 16000:
            with torch.no_grad():
                mel = audio.mel_spectrogram(y, n_fft=1920, num_mels=80, sampling_rate=24000, hop_size=480, win_size=1920, fmin=0,
                                            fmax=8000, center=False)
            out[f"{name}.mel"] = mel[0].numpy().astype(np.float32)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
