"""Timeline of one flash_attn_v9 CTA sharing its SM with three others at the bench shape (CV2_TRACE_ATTN=1: the kernel logs clock64
stamps of its first softmax warp and of its MMA warp into the debug trace buffer).

    CV2_TRACE_ATTN=1 python profiles/attn_trace.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from cosyvoice2_eu_b200 import lib

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
buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
run = lambda: lib.check(L.cv2_op_flash_attn(st, lib.ptr(q), lib.ptr(k), lib.ptr(vt), lib.ptr(out), lib.ptr(lens_d), 0, S, H, T, 0))
for _ in range(3):
    run()
torch.cuda.synchronize()
L.cv2_debug_set_ffn_trace(lib.ptr(buf))
run()
torch.cuda.synchronize()
L.cv2_debug_set_ffn_trace(None)
b = buf.cpu().numpy()
names = {1: "softmax: wait s_full", 2: "softmax: s_full seen", 3: "softmax: row max known", 4: "softmax: exps done, P stored", 5: "softmax: st wait done",
         6: "softmax: arrived p_full", 20: "mma: wait p_full", 21: "mma: p_full seen", 22: "mma: PV + S(j+1) + commits issued"}
evs = []
for off, who in ((0, "softmax warp 0"), (2048, "MMA warp")):
    evs += [(int(x) >> 8, int(x) & 255, who) for x in b[off:off + 2048] if x != 0]
evs.sort()
t0 = evs[0][0]
last = {}
for t, c, who in evs[:int(os.environ.get("N_EVENTS", "120"))]:
    print(f"  {t - t0:8d}  (+{t - last.get(who, t):6d})  {names.get(c, c)}")
    last[who] = t
