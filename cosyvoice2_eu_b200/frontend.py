"""Prompt-feature front end of token2wav on the B200, backed by `cv2_prompt_mel`.

Mirrors (names, argument meaning, return layout):
  * matcha.utils.audio.mel_spectrogram      third_party/Matcha-TTS/matcha/utils/audio.py:45-82  -> mel_spectrogram
    (at the cosyvoice2.yaml:152-160 settings, the only ones the reference's `feat_extractor` is built with)
  * CosyVoiceFrontEnd._extract_speech_feat  cosyvoice/cli/frontend.py:285-289                   -> extract_speech_feat
  * the "force speech_feat % speech_token = 2" truncation  cosyvoice/cli/frontend.py:498-502    -> align_prompt
  * torchaudio.transforms.Resample(orig_freq=16000, new_freq=24000)(prompt_speech_16k)  cosyvoice/cli/frontend.py:495,541
                                                                                               -> resample_16k_to_24k
plus `extract_speech_feat_batch` for many prompts at once (the reference handles one request at a time).

torch is used for device memory and streams only; the arithmetic runs in libcv2eu_b200.so (no CPU path: without the CUDA
library or a B200 the calls raise).
"""
import ctypes as C

import torch

from . import lib as _lib

N_FFT, NUM_MELS, SAMPLING_RATE, HOP_SIZE, WIN_SIZE, FMIN, FMAX = 1920, 80, 24000, 480, 1920, 0, 8000


def _check_config(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center):
    got = (n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, bool(center))
    want = (N_FFT, NUM_MELS, SAMPLING_RATE, HOP_SIZE, WIN_SIZE, FMIN, FMAX, False)
    if got != want:
        raise _lib.Cv2Error(f"mel_spectrogram: only the CosyVoice2 feat_extractor configuration {want} is built, got {got}")


def _run(wav, n_samples, max_samples):
    """wav [B, stride] fp32 cuda, n_samples [B] int32 cuda -> (mel [B, T, 80] fp32, mel_len [B] int32) on the device."""
    L = _lib.load()
    if not wav.is_cuda:
        raise _lib.Cv2Error("prompt features are computed on the GPU only: pass a CUDA tensor (there is no CPU fallback)")
    B = wav.shape[0]
    T = int(L.cv2_prompt_mel_frames(int(max_samples)))
    if T <= 0:
        raise _lib.Cv2Error(f"prompt of {max_samples} samples is too short: reflect padding needs more than 720")
    mel = torch.empty(B, T, NUM_MELS, dtype=torch.float32, device=wav.device)
    mel_len = torch.empty(B, dtype=torch.int32, device=wav.device)
    nbytes = int(L.cv2_prompt_mel_workspace_bytes(B, int(max_samples)))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=wav.device)
    st = C.c_void_p(torch.cuda.current_stream(wav.device).cuda_stream)
    with torch.cuda.device(wav.device):
        _lib.check(L.cv2_prompt_mel(st, _lib.ptr(wav), wav.stride(0), _lib.ptr(n_samples), B, int(max_samples), _lib.ptr(mel),
                                    _lib.ptr(mel_len), _lib.ptr(ws), nbytes))
    return mel, mel_len


def mel_spectrogram(y, n_fft=N_FFT, num_mels=NUM_MELS, sampling_rate=SAMPLING_RATE, hop_size=HOP_SIZE, win_size=WIN_SIZE,
                    fmin=FMIN, fmax=FMAX, center=False):
    """y: [B, L] waveform in [-1, 1] on the GPU -> [B, 80, T] log-mel, the reference's return layout (audio.py:82)."""
    _check_config(n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, center)
    y = y.to(torch.float32).contiguous()
    if y.dim() == 1:
        y = y[None]
    n = torch.full((y.shape[0],), y.shape[1], dtype=torch.int32, device=y.device)
    mel, _ = _run(y, n, y.shape[1])
    return mel.transpose(1, 2)


def extract_speech_feat(speech, device="cuda:0"):
    """frontend.py:285-289: speech [1, L] (24 kHz) -> (speech_feat [1, T, 80], speech_feat_len [1] int32) on `device`."""
    y = speech.to(device=device, dtype=torch.float32).contiguous()
    n = torch.full((y.shape[0],), y.shape[1], dtype=torch.int32, device=y.device)
    return _run(y, n, y.shape[1])


def extract_speech_feat_batch(speeches, device="cuda:0"):
    """A list of [L_i] / [1, L_i] prompts -> (feat [B, T_max, 80] zero-padded, feat_len [B] int32), one launch pair."""
    flat = [s.reshape(-1) for s in speeches]
    lens = [int(s.numel()) for s in flat]
    max_len = max(lens)
    host = torch.zeros(len(flat), max_len, dtype=torch.float32, pin_memory=True)
    for i, s in enumerate(flat):
        host[i, :lens[i]] = s
    wav = host.to(device, non_blocking=True)
    n = torch.tensor(lens, dtype=torch.int32).to(device, non_blocking=True)
    return _run(wav, n, max_len)


def resample_16k_to_24k(speech_16k, device="cuda:0", lengths=None):
    """frontend.py:495: speech_16k [B, L] (or [L]) at 16 kHz -> [B, ceil(3 L / 2)] at 24 kHz on `device` (sinc_interp_hann,
    lowpass_filter_width 6, rolloff 0.99: torchaudio's defaults).  `lengths` ([B] int32, optional) gives ragged rows; rows are
    zero-padded and (wav24, n_out) is returned in that case."""
    L = _lib.load()
    x = speech_16k.to(device=device, dtype=torch.float32)
    if x.dim() == 1:
        x = x[None]
    x = x.contiguous()
    if not x.is_cuda:
        raise _lib.Cv2Error("resampling runs on the GPU only (there is no CPU fallback)")
    B, max_in = x.shape
    n_in = (torch.full((B,), max_in, dtype=torch.int32, device=x.device) if lengths is None
            else lengths.to(device=x.device, dtype=torch.int32))
    max_out = int(L.cv2_resample_16k_24k_len(int(max_in)))
    y = torch.empty(B, max_out, dtype=torch.float32, device=x.device)
    n_out = torch.empty(B, dtype=torch.int32, device=x.device)
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    with torch.cuda.device(x.device):
        _lib.check(L.cv2_resample_16k_24k(st, _lib.ptr(x), x.stride(0), _lib.ptr(n_in), B, int(max_in), _lib.ptr(y), y.stride(0),
                                          _lib.ptr(n_out)))
    return y if lengths is None else (y, n_out)


def extract_spk_feat(speech_16k, device="cuda:0", lengths=None):
    """frontend.py:276-278, the feature half of _extract_spk_embedding: speech_16k [B, L] (or [L]) at 16 kHz ->
    `kaldi.fbank(speech, num_mel_bins=80, dither=0, sample_frequency=16000)` minus its mean over frames, [B, T, 80] on `device`
    (+ the frame counts [B] int32 when `lengths` gives ragged rows).  The CAM++ ONNX session that turns it into the 192-d
    x-vector stays on the reference's side."""
    L = _lib.load()
    x = speech_16k.to(device=device, dtype=torch.float32)
    if x.dim() == 1:
        x = x[None]
    x = x.contiguous()
    if not x.is_cuda:
        raise _lib.Cv2Error("x-vector features are computed on the GPU only (there is no CPU fallback)")
    B, max_n = x.shape
    if max_n < 400:
        raise _lib.Cv2Error(f"kaldi.fbank needs at least one 400-sample frame, got {max_n} samples")
    n = (torch.full((B,), max_n, dtype=torch.int32, device=x.device) if lengths is None
         else lengths.to(device=x.device, dtype=torch.int32))
    T = int(L.cv2_kaldi_fbank_frames(int(max_n)))
    feat = torch.empty(B, T, NUM_MELS, dtype=torch.float32, device=x.device)
    feat_len = torch.empty(B, dtype=torch.int32, device=x.device)
    st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    with torch.cuda.device(x.device):
        _lib.check(L.cv2_kaldi_fbank(st, _lib.ptr(x), x.stride(0), _lib.ptr(n), B, int(max_n), _lib.ptr(feat), _lib.ptr(feat_len), 1))
    return feat if lengths is None else (feat, feat_len)


def align_prompt(speech_feat, speech_feat_len, speech_token, speech_token_len):
    """frontend.py:498-502 (CosyVoice2, 24 kHz): keep n = min(feat frames // 2, tokens) tokens and exactly 2 n mel frames.
    The two length tensors are updated in place, as the reference does."""
    n_tok = min(speech_feat.shape[1] // 2, speech_token.shape[1])
    speech_feat_len.fill_(2 * n_tok)
    speech_token_len.fill_(n_tok)
    return speech_feat[:, :2 * n_tok], speech_feat_len, speech_token[:, :n_tok], speech_token_len
