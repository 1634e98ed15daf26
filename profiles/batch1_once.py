"""One eager batch-1 token2wav (configs[1]: 250 tokens + 75 prompt) for an ncu launch list:
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/batch1_once.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosyvoice2_eu_b200 import B200Flow, B200HiFT, B200Token2Wav  # noqa: E402
from synth import weights  # noqa: E402

flow, hift = B200Flow("cuda:0"), B200HiFT("cuda:0")
flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
hift.load_state_dict(weights.to_torch(weights.make_hift_state()))
t2w = B200Token2Wav(flow, hift)
u = weights.make_utterance(250, 75, seed=99)
a = [torch.from_numpy(u[k][0]) for k in ("token", "prompt_token", "prompt_feat", "embedding")]
for i in range(3):
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    t2w.token2wav_batch([a[0]], [a[1]], [a[2]], [a[3]])
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
