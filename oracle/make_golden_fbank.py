"""tests/golden/fbank.npz: the x-vector features exactly as CosyVoiceFrontEnd._extract_spk_embedding computes them
(cosyvoice/cli/frontend.py:276-278) -- torchaudio.compliance.kaldi.fbank(speech, num_mel_bins=80, dither=0,
sample_frequency=16000) minus its column mean -- on seeded 16 kHz prompts: a voiced glide, noise, a ragged length, the
400-sample minimum (one frame), digital silence (the epsilon floor) and a loud clipped one.
    python oracle/make_golden_fbank.py        (needs torchaudio; no /root/reference code is involved)
"""
import os

import numpy as np
import torch
import torchaudio.compliance.kaldi as kaldi

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def prompts():
    rng = np.random.Generator(np.random.Philox(key=1600))
    t = np.arange(16000 * 3) / 16000.0
    f0 = 110 + 60 * t
    glide = 0.3 * np.sin(2 * np.pi * np.cumsum(f0) / 16000.0) * (1 + 0.3 * np.sin(2 * np.pi * 3 * t))
    return {
        "glide": glide.astype(np.float32),
        "noise": (0.1 * rng.standard_normal(16000 * 2 + 77)).astype(np.float32),
        "ragged": (0.2 * rng.standard_normal(5000) * np.linspace(0, 1, 5000)).astype(np.float32),
        "one_frame": (0.05 * rng.standard_normal(400)).astype(np.float32),
        "silence": np.zeros(1234, np.float32),
        "clipped": np.clip(3.0 * rng.standard_normal(8000), -1, 1).astype(np.float32),
    }


def main():
    d = {}
    for name, w in prompts().items():
        feat = kaldi.fbank(torch.from_numpy(w)[None], num_mel_bins=80, dither=0, sample_frequency=16000)
        raw = feat.numpy().copy()
        feat = feat - feat.mean(dim=0, keepdim=True)
        d[f"{name}.wav"] = w
        d[f"{name}.fbank_raw"] = raw
        d[f"{name}.feat"] = feat.numpy()
        print(name, w.shape, "->", tuple(feat.shape), "raw range %.2f .. %.2f" % (raw.min(), raw.max()))
    np.savez_compressed(os.path.join(OUT, "fbank.npz"), **d)


if __name__ == "__main__":
    main()
