"""Does running the bench step as TWO concurrent half batches (two host threads, two CUDA streams, one engine) recover the wave
tails of the persistent kernels (FFN: 319 pair units over 74 CTA pairs = 4.31 rounds run as 5)?

    python profiles/two_stream_step.py [steps]

Prints ms per step for the whole 64-utterance batch on one stream and for two interleaved halves on two streams."""
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosyvoice2_eu_b200 import B200Flow, B200HiFT, B200Token2Wav  # noqa: E402
from synth import weights  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
dev = "cuda:0"
flow, hift = B200Flow(dev), B200HiFT(dev)
flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
hift.load_state_dict(weights.to_torch(weights.make_hift_state()))
t2w = B200Token2Wav(flow, hift)
rng = np.random.Generator(np.random.Philox(key=1000))
n_tokens = sorted(int(round(25 * d)) for d in rng.uniform(4.0, 20.0, size=64))
utts = [weights.make_utterance(n, 75, seed=i) for i, n in enumerate(n_tokens)]
cols = lambda idx: tuple([torch.from_numpy(utts[i][k][0]) for i in idx] for k in ("token", "prompt_token", "prompt_feat", "embedding"))
whole = cols(range(64))
halves = [cols(range(0, 64, 2)), cols(range(1, 64, 2))]
audio_s = sum(2 * n * 480 for n in n_tokens) / 24000.0


def run_whole():
    speech, _ = t2w.token2wav_batch(*whole)
    return speech


streams = [torch.cuda.Stream(), torch.cuda.Stream()]


def run_halves():
    out = [None, None]

    def work(i):
        with torch.cuda.stream(streams[i]):
            out[i], _ = t2w.token2wav_batch(*halves[i])

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    return out


for name, fn in (("whole batch, one stream", run_whole), ("two half batches, two streams", run_halves), ("whole batch, one stream", run_whole)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    print(f"{name}: {ms:.1f} ms per step, {audio_s / ms * 1e3:.1f} audio-s/s")
