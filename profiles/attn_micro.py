"""Micro-benchmark of the estimator attention kernel at the bench's shape (configs[2]: 64 utterances x 2 CFG rows,
8 heads, mel lengths 2*(U[100,500]+75)), timed with CUDA events through the C ABI (cv2_op_flash_attn).

    python profiles/attn_micro.py [reps] [chunk]       (CV2_ATTN_V6=1 selects the previous kernel)
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from cosyvoice2_eu_b200 import lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = lib.load()
rng = np.random.Generator(np.random.Philox(key=1000))
n_tok = sorted(int(round(25 * d)) for d in rng.uniform(4.0, 20.0, size=64))
lens = [2 * (n + 75) for n in n_tok] * 2
S, H, D = len(lens), 8, 64
T = (max(lens) + 127) // 128 * 128
g = torch.Generator(device="cuda").manual_seed(1)
q = (torch.randn(S, H, T, D, generator=g, device="cuda") * 0.25).half()
k = (torch.randn(S, H, T, D, generator=g, device="cuda") * 2).half()
vt = torch.randn(S, H, D, T, generator=g, device="cuda").half()
out = torch.zeros(S, T, H * D, dtype=torch.float16, device="cuda")
lens_d = torch.tensor(lens, dtype=torch.int32, device="cuda")
st = torch.cuda.current_stream().cuda_stream


def run():
    lib.check(L.cv2_op_flash_attn(st, lib.ptr(q), lib.ptr(k), lib.ptr(vt), lib.ptr(out), lib.ptr(lens_d), 0, S, H, T, chunk))


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flop = sum(4.0 * 512 * n * n for n in lens) if chunk == 0 else sum(4.0 * 512 * n * (n + chunk) / 2 for n in lens)
print(f"flash_attn S={S} T={T} chunk={chunk}: {ms * 1e3:.1f} us/launch, {flop / ms / 1e9:.1f} TFLOP/s (algorithmic)")
