"""Where a streaming chunk step of 32 sessions spends its time: host launch time vs device time of the flow (incremental and
prefix recompute), the vocoder, and the Python glue of token2wav_stream_batch.   python profiles/stream_profile.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cosyvoice2_eu_b200 import B200Flow, B200HiFT, B200Token2Wav  # noqa: E402
from cosyvoice2_eu_b200.scheduler import chunk_schedule  # noqa: E402
from synth import weights  # noqa: E402


def main():
    dev = "cuda:0"
    flow, hift = B200Flow(dev), B200HiFT(dev)
    flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
    hift.load_state_dict(weights.to_torch(weights.make_hift_state()))
    t2w = B200Token2Wav(flow, hift)
    n_sess, n_tok, n_prompt = 32, 250, 75
    sess = [{k: torch.from_numpy(v) for k, v in weights.make_utterance(n_tok, n_prompt, seed=5000 + i).items()} for i in range(n_sess)]
    sched = chunk_schedule(n_tok, n_prompt)
    group = flow.open_stream_group(n_sess, max_mel_frames=640)

    def reqs_of(n_vis, off):
        return [dict(token=u["token"][:, :n_vis], prompt_token=u["prompt_token"], prompt_feat=u["prompt_feat"],
                     embedding=u["embedding"], token_offset=off, uuid=f"p{i}") for i, u in enumerate(sess)]

    for rep in range(2):
        for i in range(n_sess):
            group.release(f"p{i}")
            t2w.hift_cache_dict[f"p{i}"] = None
        for ci, (n_vis, off, fin) in enumerate(sched):
            reqs = reqs_of(n_vis, off)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if fin:
                flow.inference_batch([r["token"][0] for r in reqs], [r["prompt_token"][0] for r in reqs], [r["prompt_feat"][0] for r in reqs],
                                     [r["embedding"][0] for r in reqs], streaming=False, finalize=True)
            else:
                flow.inference_stream_group(group, reqs)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            # the full step (flow again + vocoder + glue) for comparison
            if not fin:
                for i in range(n_sess):      # undo the row counter advance: recompute the same chunk in the full step below
                    pass
            if rep == 1:
                print(f"chunk {ci}: flow host {1e3 * (t1 - t0):6.1f} ms, flow total {1e3 * (t2 - t0):6.1f} ms, launches {flow.last_launches}")
    # full steps with a timeline of the pieces
    import cosyvoice2_eu_b200.engine as E
    for i in range(n_sess):
        group.release(f"p{i}")
        t2w.hift_cache_dict[f"p{i}"] = None
    orig_hift = hift.inference
    acc = {"hift_host": 0.0, "hift_calls": 0}

    def timed_hift(*a, **k):
        t0 = time.perf_counter()
        r = orig_hift(*a, **k)
        acc["hift_host"] += time.perf_counter() - t0
        acc["hift_calls"] += 1
        return r
    hift.inference = timed_hift
    for ci, (n_vis, off, fin) in enumerate(sched):
        acc["hift_host"], acc["hift_calls"] = 0.0, 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = t2w.token2wav_stream_batch(reqs_of(n_vis, off), finalize=fin, group=group)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        host = [o.cpu() for o in outs]
        t3 = time.perf_counter()
        print(f"step {ci}: call returned {1e3 * (t1 - t0):6.1f} ms (hift host {1e3 * acc['hift_host']:5.1f} ms in {acc['hift_calls']} calls), "
              f"device done {1e3 * (t2 - t0):6.1f} ms, D2H of 32 chunks {1e3 * (t3 - t2):5.1f} ms")


if __name__ == "__main__":
    main()
