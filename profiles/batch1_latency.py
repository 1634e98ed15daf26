import sys, time, os, numpy as np, torch
sys.path.insert(0, '/root/repo')
from cosyvoice2_eu_b200 import B200Flow, B200HiFT, B200Token2Wav, GraphedToken2Wav
from synth import weights
flow, hift = B200Flow("cuda:0"), B200HiFT("cuda:0")
flow.load_state_dict(weights.to_torch(weights.make_flow_state())); hift.load_state_dict(weights.to_torch(weights.make_hift_state()))
t2w = B200Token2Wav(flow, hift)
u1 = weights.make_utterance(250, 75, seed=99)
a1 = [torch.from_numpy(u1[k][0]) for k in ("token", "prompt_token", "prompt_feat", "embedding")]
g1 = GraphedToken2Wav(t2w)
for _ in range(3): g1([a1[0]], [a1[1]], [a1[2]], [a1[3]])
torch.cuda.synchronize()
lat = []
for _ in range(9):
    t0 = time.perf_counter(); w, _ = g1([a1[0]], [a1[1]], [a1[2]], [a1[3]]); w.cpu(); lat.append(time.perf_counter() - t0)
print("min_2sm_tiles", os.environ.get("CV2_MIN_2SM_TILES"), "batch-1 latency ms", 1e3 * float(np.median(lat)))
