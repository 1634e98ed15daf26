"""One incremental streaming chunk of 32 sessions (chunk index 4: 406 mel frames, 2 tiles per sequence at most) for an ncu launch
list:  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python profiles/stream_chunk_once.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosyvoice2_eu_b200 import B200Flow  # noqa: E402
from cosyvoice2_eu_b200.scheduler import chunk_schedule  # noqa: E402
from synth import weights  # noqa: E402

flow = B200Flow("cuda:0")
flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
n_sess = int(os.environ.get("N_SESS", "32"))
sess = [{k: torch.from_numpy(v) for k, v in weights.make_utterance(250, 75, seed=5000 + i).items()} for i in range(n_sess)]
sched = chunk_schedule(250, 75)
group = flow.open_stream_group(n_sess, max_mel_frames=640)
upto = int(os.environ.get("CHUNK", "5"))
for ci, (n_vis, off, fin) in enumerate(sched[:upto + 1]):
    reqs = [dict(token=u["token"][:, :n_vis], prompt_token=u["prompt_token"], prompt_feat=u["prompt_feat"], embedding=u["embedding"],
                 token_offset=off, uuid=f"p{i}") for i, u in enumerate(sess)]
    if ci == upto:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    flow.inference_stream_group(group, reqs)
    torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
