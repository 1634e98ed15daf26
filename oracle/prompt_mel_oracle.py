"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy, float64 arithmetic, float32 result) of the prompt-feature front end that feeds token2wav
(SURVEY.md section 8f row F2): the 24 kHz log-mel spectrogram of the prompt waveform,

    MT/utils/audio.py:45-82        mel_spectrogram(y, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center)
    cosyvoice2.yaml:152-160        n_fft 1920, num_mels 80, sampling_rate 24000, hop 480, win 1920, fmin 0, fmax 8000, center False
    CV/cli/frontend.py:285-289     _extract_speech_feat: [1, 80, T] -> [1, T, 80] + length
    CV/cli/frontend.py:498-502     "force speech_feat % speech_token = 2" truncation (integer bookkeeping)

Pinning: tests/golden/prompt_mel.npz holds outputs of the reference's own `mel_spectrogram` (imported unmodified from
/root/reference by oracle/make_golden_prompt_mel.py) and tests/test_oracle_golden.py checks this file against them.  The mel
filterbank itself comes from `librosa.filters.mel` (librosa is a requirements.txt dependency that is absent from
/root/reference and from this image): restated below from its published algorithm (Slaney scale, Slaney area norm) and
pinned against the independent implementation that IS in the image, `transformers.audio_utils.mel_filter_bank(norm="slaney",
mel_scale="slaney")`, which the golden generator also hands to the reference in place of the missing import.  The reference
holds no test of its own for this function, so the mel basis is "pinned against a third-party restatement", the STFT / log
part against the reference itself.

`resample_16k_to_24k` restates `torchaudio.transforms.Resample(orig_freq=16000, new_freq=24000)` (the call at
CV/cli/frontend.py:495,541; torchaudio is a requirements.txt dependency outside /root/reference, version 2.11 in this image:
sinc interpolation with a Hann window, lowpass_filter_width 6, rolloff 0.99) and is pinned against torchaudio itself
(`resample.*` entries of the golden file, written by the same generator).

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

N_FFT = 1920
HOP = 480
N_MELS = 80
SR = 24000
FMIN = 0.0
FMAX = 8000.0
PAD = (N_FFT - HOP) // 2          # MT/utils/audio.py:57-59
N_BINS = N_FFT // 2 + 1


def _hz_to_mel(f):
    """librosa.convert.hz_to_mel, htk=False (Slaney): linear below 1 kHz, logarithmic above."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    big = f >= min_log_hz
    return np.where(big, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_basis(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney', float32):
    triangles between n_mels + 2 points equally spaced on the Slaney mel scale, each scaled to unit area."""
    fftfreqs = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    w = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    w *= enorm[:, None]                    # float32 array scaled in place, as librosa does
    return w


def hann_periodic(n=N_FFT):
    """torch.hann_window(n) (periodic=True): 0.5 - 0.5 cos(2 pi k / n)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def n_frames(n_samples):
    """Frames of the centre=False STFT over the reflect-padded signal (MT/utils/audio.py:57-75)."""
    return (n_samples + 2 * PAD - N_FFT) // HOP + 1


def mel_spectrogram(y):
    """y: [L] or [B, L] waveform in [-1, 1] -> [B, 80, T] float32 log-mel (the reference's return layout)."""
    y = np.asarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y[None]
    assert y.shape[1] > PAD, "reflect padding needs more than 720 samples"
    yp = np.pad(y, ((0, 0), (PAD, PAD)), mode="reflect")                 # audio.py:57-60
    T = (yp.shape[1] - N_FFT) // HOP + 1
    idx = np.arange(T)[:, None] * HOP + np.arange(N_FFT)[None, :]
    frames = yp[:, idx] * hann_periodic()[None, None, :]                 # audio.py:62-75 (torch.stft, center=False)
    spec = np.fft.rfft(frames, axis=-1)                                  # [B, T, 961]
    mag = np.sqrt(spec.real ** 2 + spec.imag ** 2 + 1e-9)                # audio.py:77
    mel = mag @ mel_basis().astype(np.float64).T                         # audio.py:79
    out = np.log(np.maximum(mel, 1e-5))                                  # audio.py:80 -> :22-23
    return out.transpose(0, 2, 1).astype(np.float32)


def extract_speech_feat(speech):
    """CV/cli/frontend.py:285-289: [1, L] -> ([1, T, 80] float32, [1] int32)."""
    feat = mel_spectrogram(speech).transpose(0, 2, 1)
    return np.ascontiguousarray(feat), np.array([feat.shape[1]], dtype=np.int32)


def align_prompt(feat_len, token_len):
    """CV/cli/frontend.py:498-502 (integer, bit-exact): token_len' = min(feat_len // 2, token_len), feat_len' = 2 token_len'."""
    t = min(int(feat_len) // 2, int(token_len))
    return 2 * t, t


# ------------------------------------------------------------------------------------------ 16 kHz -> 24 kHz
RS_ORIG, RS_NEW, RS_LPW, RS_ROLLOFF = 2, 3, 6, 0.99      # 16000 / 24000 reduced by their gcd; torchaudio defaults
RS_WIDTH = int(np.ceil(RS_LPW * RS_ORIG / (min(RS_ORIG, RS_NEW) * RS_ROLLOFF)))          # 7
RS_TAPS = 2 * RS_WIDTH + RS_ORIG                                                           # 16


def resample_kernel():
    """torchaudio.functional._get_sinc_resample_kernel(16000, 24000, gcd 8000): [3 phases, 16 taps] float32.
    Phase p's filter is sinc(pi t) cos^2(pi t / 12) * 0.99 at t = 1.98 (k - 7) / 2 - 1.98 p / 3, t clamped to [-6, 6]
    (the phase offset is formed in float32 first, as torch's int / int division does, then everything runs in float64)."""
    base = min(RS_ORIG, RS_NEW) * RS_ROLLOFF
    idx = np.arange(-RS_WIDTH, RS_WIDTH + RS_ORIG, dtype=np.float64)[None, :] / RS_ORIG
    ph = (np.arange(0, -RS_NEW, -1).astype(np.float32) / np.float32(RS_NEW)).astype(np.float64)[:, None]
    t = np.clip((ph + idx) * base, -RS_LPW, RS_LPW)
    window = np.cos(t * np.pi / RS_LPW / 2) ** 2
    t = t * np.pi
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(t == 0, 1.0, np.sin(t) / t)
    return (k * window * (base / RS_ORIG)).astype(np.float32)


def resample_len(n_in):
    """ceil(3 n / 2) (functional._apply_sinc_resample_kernel: target_length)."""
    return (RS_NEW * int(n_in) + RS_ORIG - 1) // RS_ORIG


def resample_16k_to_24k(x):
    """x: [L] or [B, L] -> [B, ceil(3L/2)] float32: y[3 i + p] = sum_k xpad[2 i + k] h[p][k], xpad = x padded 7 left, 9 right."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[None]
    L = x.shape[1]
    h = resample_kernel().astype(np.float64)
    xp = np.pad(x, ((0, 0), (RS_WIDTH, RS_WIDTH + RS_ORIG)))
    n_i = (xp.shape[1] - RS_TAPS) // RS_ORIG + 1
    idx = np.arange(n_i)[:, None] * RS_ORIG + np.arange(RS_TAPS)[None, :]
    y = np.einsum("bik,pk->bip", xp[:, idx], h).reshape(x.shape[0], -1)
    return y[:, :resample_len(L)].astype(np.float32)
