import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from synth import weights
from cosyvoice2_eu_b200 import B200Flow, lib

flow = B200Flow("cuda:0")
flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
rng = np.random.Generator(np.random.Philox(key=1000))
n_tok = sorted(int(round(25 * d)) for d in rng.uniform(4.0, 20.0, size=64))
us = [weights.make_utterance(n, 75, seed=i) for i, n in enumerate(n_tok)]
args = [[torch.from_numpy(u[k][0]) for u in us] for k in ("token", "prompt_token", "prompt_feat", "embedding")]
flow.inference_batch(*args)
torch.cuda.synchronize()
buf = torch.zeros(16384, dtype=torch.int64, device="cuda")
L = lib.load()
L.cv2_debug_set_ffn_trace(lib.ptr(buf))
flow.inference_batch(*args)
torch.cuda.synchronize()
L.cv2_debug_set_ffn_trace(None)
b = buf.cpu().numpy()[8192:]
names = {1: "mma: tile start", 2: "mma: tmem_empty ok", 3: "mma: stage full", 4: "mma: tile committed",
         30: "epi: q chunk store", 31: "epi: k chunk store", 32: "epi: v chunk store", 10: "epi: wait tmem_full", 11: "epi: tmem_full ok", 12: "epi: chunk tmem_ld+bias done", 13: "epi: residual added", 14: "epi: out32 stored",
         15: "epi: chunks done", 16: "epi: pre bar1", 17: "epi: bar1 ok", 18: "epi: pre bar2", 19: "epi: bar2 ok", 20: "epi: tile done"}
allev = []
for off, who in ((0, "epi warp 0"), (2048, "epi warp 13"), (4096, "MMA thread")):
    ev = [(int(x) >> 8, int(x) & 255) for x in b[off:off + (2048 if off < 4096 else 4096)] if x != 0]
    allev.append((who, ev))
t0 = min(ev[0][0] for _, ev in allev if ev)
for who, ev in allev:
    print(f"== {who}: {len(ev)} events")
    last = ev[0][0] if ev else 0
    for t, c in ev[:150]:
        print(f"  {t - t0:8d}  (+{t - last:6d})  {names.get(c, c)}")
        last = t
