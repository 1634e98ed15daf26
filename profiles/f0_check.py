import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
from synth import weights
from cosyvoice2_eu_b200 import B200HiFT
hs = weights.to_torch(weights.make_hift_state())
hift = B200HiFT("cuda:0"); hift.load_state_dict(hs)
g = np.load('tests/golden/tiny.npz')
noise = torch.from_numpy(weights.make_nsf_noise(g["mel"].shape[2] * 480, int(g["seed"])))
sp, so, f0 = hift.inference(torch.from_numpy(g["mel"]), noise=noise, return_f0=True)
e = np.abs(f0.cpu().numpy() - g["f0"])
print(os.environ.get("CV2_F0_FP32"), "f0 max err", e.max(), "mean", e.mean(), "argmax", e.argmax(), "f0 there", g["f0"].reshape(-1)[e.argmax()])
