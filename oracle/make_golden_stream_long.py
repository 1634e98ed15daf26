"""tests/golden/stream_long.npz: the flow mel of every chunk call of a LONGER streaming session from the UNMODIFIED reference
(container only, same fixtures / shims as make_golden.py): P = 75 prompt tokens, N = 160 tokens, hop 25 -> non-final chunk calls
of 200, 250, 300, 350, 400, 450 mel frames (they cross the 128 / 256 / 384-row tile boundaries the incremental engine resumes
from) and the final full-attention call.  Stored per chunk: the frames token2wav keeps (from 2 * token_offset on,
CV/cli/model.py:311).      python oracle/make_golden_stream_long.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import weights  # noqa: E402
from make_golden import OUT, flow_call  # noqa: E402
from token2wav_oracle import stream_schedule  # noqa: E402


def main():
    torch.manual_seed(0)
    flow, _ = ref_shims.build_reference_modules()
    flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
    n_tok, n_prompt, seed = 160, 75, 33
    u = weights.make_utterance(n_tok, n_prompt, seed)
    sched = stream_schedule(n_tok, n_prompt)
    d = dict(n_tok=n_tok, n_prompt=n_prompt, seed=seed, schedule=np.array([(a, b, int(c)) for a, b, c in sched], np.int64))
    with torch.inference_mode():
        for ci, (n_vis, off, fin) in enumerate(sched):
            uu = dict(u, token=u["token"][:, :n_vis])
            mel = flow_call(flow, uu, streaming=not fin, finalize=fin)
            d[f"mel{ci}"] = mel[:, :, 2 * off:].numpy().copy()
            d[f"mel_len{ci}"] = mel.shape[2]
            print("chunk", ci, (n_vis, off, fin), "mel", tuple(mel.shape), "kept", d[f"mel{ci}"].shape)
    np.savez_compressed(os.path.join(OUT, "stream_long.npz"), **d)


if __name__ == "__main__":
    main()
