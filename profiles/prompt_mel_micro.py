"""Prompt log-mel front end (cv2_prompt_mel) timed with CUDA events, the oracle (numpy, float64 FFT) beside it.

    python profiles/prompt_mel_micro.py [B] [seconds] [reps]

Algorithmic work per frame: 2 x 961 x 961 fp32 FMAs (folded real DFT) = 3.69 MFLOP; bytes: 480 new samples read, 80 floats
written (the 961-bin magnitude row makes one 4 KB round trip through L2/HBM between the two kernels)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
from cosyvoice2_eu_b200 import extract_speech_feat_batch, frontend

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 30.0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
L = int(24000 * secs)
g = torch.Generator().manual_seed(3)
waves = [(torch.rand(L, generator=g) * 2 - 1) * 0.5 for _ in range(B)]
wav = torch.stack(waves).cuda()
n = torch.full((B,), L, dtype=torch.int32, device="cuda")
for _ in range(3):
    mel, mel_len = frontend._run(wav, n, L)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    mel, mel_len = frontend._run(wav, n, L)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
frames = int(mel_len.sum())
print(f"cv2_prompt_mel B={B} x {secs:.0f} s: {ms:.3f} ms/call, {frames} frames, {frames * 3.694e6 / ms / 1e9:.1f} TFLOP/s fp32 (algorithmic), "
      f"{B * secs / (ms / 1e3):.0f} audio-s/s")
t0 = time.perf_counter()
for _ in range(reps):
    mel, mel_len = extract_speech_feat_batch(waves)
    mel_len.cpu()
t1 = time.perf_counter()
print(f"host buffers in, lengths out (H2D of {B * L * 4 / 1e6:.1f} MB inside): {(t1 - t0) / reps * 1e3:.3f} ms/call")
import prompt_mel_oracle as PO
w0 = waves[0].numpy()
t0 = time.perf_counter()
ref = PO.mel_spectrogram(w0)
t1 = time.perf_counter()
err = np.abs(mel[0].cpu().numpy() - ref[0].T).max()
print(f"oracle (numpy float64 rfft, 1 thread) one {secs:.0f} s prompt: {(t1 - t0) * 1e3:.1f} ms -> {secs / (t1 - t0):.0f} audio-s/s; max |log-mel diff| = {err:.2e}")
