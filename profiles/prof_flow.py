"""Profiling driver: one flow.inference + hift.inference at a chosen batch (used under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from synth import weights
from cosyvoice2_eu_b200 import B200Flow, B200HiFT, B200Token2Wav

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n_tok = int(sys.argv[2]) if len(sys.argv) > 2 else 250
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
flow, hift = B200Flow("cuda:0"), B200HiFT("cuda:0")
flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
hift.load_state_dict(weights.to_torch(weights.make_hift_state()))
t2w = B200Token2Wav(flow, hift)
us = [weights.make_utterance(n_tok - 7 * i, 75, seed=i) for i in range(B)]
args = [[torch.from_numpy(u[k][0]) for u in us] for k in ("token", "prompt_token", "prompt_feat", "embedding")]
for _ in range(reps):
    w, l = t2w.token2wav_batch(*args)
torch.cuda.synchronize()
print("ok", tuple(w.shape))
