"""Timeline of one CTA of ffn_fused_kernel (cv2_debug_set_ffn_trace): where a tile's ~70 k cycles go.
    python profiles/ffn_trace.py"""
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
buf = torch.zeros(8192, dtype=torch.int64, device="cuda")
L = lib.load()
L.cv2_debug_set_ffn_trace(lib.ptr(buf))
flow.inference_batch(*args)          # every ffn launch overwrites the log: the last launch's timeline survives
torch.cuda.synchronize()
L.cv2_debug_set_ffn_trace(None)
b = buf.cpu().numpy()
names = {7: "  slot: wait w_full", 8: "  slot: w_full ok", 9: "  slot: 8 MMAs issued", 1: "tile start (mma)", 2: "H tile landed", 3: "FF1 issued", 4: "wait f_full", 5: "f_full seen", 6: "FF2 issued",
         10: "epi: wait acc1", 11: "epi: acc1 ready", 12: "epi: loaded+barrier", 13: "epi: gelu+st done", 14: "epi: arrived",
         20: "epi: chunks done", 21: "epi: acc2 ready", 22: "epi: tile done",
         23: "  out: acc chunk + b2 in regs", 24: "  out: residual landed + added", 25: "  out: x32 stores issued", 26: "  out: both chunks done",
         27: "  out: LayerNorm barrier passed", 30: "mid: wait op_full", 31: "mid: op_full + h_empty seen", 32: "mid: x written back, stats ready",
         33: "mid: LayerNorm barrier passed", 34: "mid: H tile built, arrived"}
for off, who in ((0, "MMA thread"), (4096, "epilogue warp 2")):
    ev = [(int(x) >> 8, int(x) & 255) for x in b[off:off + 4096] if x != 0]
    if not ev:
        continue
    t0 = ev[0][0]
    print(f"== {who}: {len(ev)} events")
    last = t0
    for t, c in ev[:int(os.environ.get("N_EVENTS", "200"))]:
        print(f"  {t - t0:8d}  (+{t - last:6d})  {names.get(c, c)}")
        last = t
