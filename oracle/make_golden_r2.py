"""Second set of golden vectors from the UNMODIFIED reference (container only; same fixtures and shims as make_golden.py).

    python oracle/make_golden_r2.py            # needs /root/reference

Files
  stream_stages.npz : the streaming session of stream.npz (P=10, N=70, hop 25) with every stage of every chunk recorded by
                      wrapping (not modifying) the reference's own callables while CosyVoice2Model.token2wav runs
                      (CV/cli/model.py:300-334): per chunk i
                        flow_mel{i}     what flow.inference returned (before the token_offset slice)
                        hift_mel{i}     the mel handed to hift.inference (8 cached frames + new frames, model.py:313-316)
                        cache_source{i} the cache_source handed to hift.inference ([1,1,0] for chunk 0, generator.py:578-580)
                        speech_pre{i}   hift.inference's speech, before fade_in_out
                        source{i}       hift.inference's source (after the cache_source overwrite)
                        fade_old{i}     the cached speech tail handed to fade_in_out (absent for chunk 0)
                        speech_post{i}  fade_in_out's result (CV/utils/common.py:142-150)
                        out{i}          what token2wav returned
  long.npz          : BASELINE configs[1] (250 tokens + 75 prompt, T = 650) and the longest utterance of configs[2]
                      (500 tokens + 75 prompt, T = 1150): flow mel, f0, hift wav (same noise injection as make_golden.py)
  tts_schedule.npz  : the (n_visible, token_offset, finalize) sequence of token2wav calls made by the reference's OWN loop,
                      CosyVoice2Model.tts(source_speech_token=..., stream=True) (model.py:336-398; vc_job hands over all tokens
                      at once, model.py:141-143), for a grid of (n_tokens, n_prompt).  With every token already there the
                      loop's break test (model.py:369) is evaluated with the hop of the chunk just emitted
                      (this_token_hop_len is not recomputed after token_offset moves), so with prompt_token_pad > 0 it can
                      finalize one chunk earlier than a slowly fed session would.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import weights  # noqa: E402
from make_golden import OUT, flow_call, inject_noise, t  # noqa: E402
from token2wav_oracle import stream_schedule  # noqa: E402


def main():
    torch.manual_seed(0)
    flow, hift = ref_shims.build_reference_modules()
    flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
    hift.load_state_dict(weights.to_torch(weights.make_hift_state()))
    model = ref_shims.build_reference_model(flow, hift)
    model.device = torch.device("cpu")
    import cosyvoice.cli.model as ref_model_mod

    # ---- streaming, stage by stage -------------------------------------------------------------
    n_tok, n_prompt, seed = 70, 10, 3
    u = weights.make_utterance(n_tok, n_prompt, seed)
    sched = stream_schedule(n_tok, n_prompt)
    rec = {}
    orig_flow_inf, orig_hift_inf, orig_fade = flow.inference, hift.inference, ref_model_mod.fade_in_out
    state = {"ci": 0}

    def flow_inf(**kw):
        mel, x = orig_flow_inf(**kw)
        rec[f"flow_mel{state['ci']}"] = mel.numpy().copy()
        return mel, x

    def hift_inf(speech_feat, cache_source=torch.zeros(1, 1, 0)):
        rec[f"hift_mel{state['ci']}"] = speech_feat.numpy().copy()
        rec[f"cache_source{state['ci']}"] = cache_source.numpy().copy()
        speech, source = orig_hift_inf(speech_feat=speech_feat, cache_source=cache_source)
        rec[f"speech_pre{state['ci']}"] = speech.numpy().copy()
        rec[f"source{state['ci']}"] = source.numpy().copy()
        return speech, source

    def fade(new, old, window):
        rec[f"fade_old{state['ci']}"] = old.numpy().copy()
        out = orig_fade(new, old, window)
        rec[f"speech_post{state['ci']}"] = out.numpy().copy()
        return out

    flow.inference, hift.inference, ref_model_mod.fade_in_out = flow_inf, hift_inf, fade
    try:
        model.hift_cache_dict["s"] = None
        with torch.inference_mode():
            for ci, (n_vis, off, fin) in enumerate(sched):
                state["ci"] = ci
                n_new = (n_vis - (0 if fin else 3)) * 2 - off * 2
                mel_len = n_new + (8 if model.hift_cache_dict["s"] is not None else 0)
                noise = t(weights.make_nsf_noise(mel_len * 480, seed * 100 + ci))
                with inject_noise([noise]):
                    w = model.token2wav(t(u["token"][:, :n_vis]), t(u["prompt_token"]), t(u["prompt_feat"]), t(u["embedding"]),
                                        token_offset=off, uuid="s", stream=not fin, finalize=fin)
                rec[f"out{ci}"] = w.numpy().copy()
    finally:
        flow.inference, hift.inference, ref_model_mod.fade_in_out = orig_flow_inf, orig_hift_inf, orig_fade
    old = dict(np.load(os.path.join(OUT, "stream.npz")))
    for ci in range(len(sched)):
        assert np.array_equal(old[f"chunk{ci}"], rec[f"out{ci}"]), ci      # same session as stream.npz, bit for bit
    rec.update(n_tok=n_tok, n_prompt=n_prompt, seed=seed, schedule=np.array([(a, b, int(c)) for a, b, c in sched], np.int64))
    np.savez_compressed(os.path.join(OUT, "stream_stages.npz"), **rec)
    print("stream_stages", sorted(rec)[:8], "...", len(rec), "entries")

    # ---- long sequences ---------------------------------------------------------------------------
    d = {}
    for name, (n_tok, n_prompt, seed) in {"cfg2": (250, 75, 21), "max": (500, 75, 22)}.items():
        uu = weights.make_utterance(n_tok, n_prompt, seed)
        with torch.inference_mode():
            mel = flow_call(flow, uu)
            noise = t(weights.make_nsf_noise(mel.shape[2] * 480, seed))
            f0 = hift.f0_predictor(mel)
            with inject_noise([noise]):
                wav, _ = hift.inference(speech_feat=mel)
        d.update({f"{name}_n_tok": n_tok, f"{name}_n_prompt": n_prompt, f"{name}_seed": seed, f"{name}_mel": mel.numpy(),
                  f"{name}_f0": f0.numpy(), f"{name}_wav": wav.numpy()})
        print(name, "mel", tuple(mel.shape), "std %.3f" % mel.std(), "wav absmax %.3f" % wav.abs().max())
    np.savez_compressed(os.path.join(OUT, "long.npz"), **d)

    # ---- the reference's own chunk loop (tts with vc_job: all tokens present from the start) ---------------
    calls = []

    def fake_token2wav(token, prompt_token, prompt_feat, embedding, token_offset, uuid, stream=False, finalize=False, speed=1.0):
        calls.append((int(token.shape[1]), int(token_offset), int(finalize)))
        return torch.zeros(1, 1)

    model.token2wav = fake_token2wav
    import time as _time
    real_sleep = _time.sleep
    ref_model_mod.time.sleep = lambda s: real_sleep(0.002)        # the 0.1 s poll quantum only slows the generator down
    out = {}
    grid = [(70, 10), (70, 25), (250, 75), (43, 10), (42, 10), (68, 10), (28, 0), (27, 0), (3, 7), (120, 60), (200, 88), (56, 49)]
    try:
        for n_tok, n_prompt in grid:
            calls.clear()
            uu = weights.make_utterance(n_tok, n_prompt, 1)
            for _ in model.tts(flow_embedding=t(uu["embedding"]), flow_prompt_speech_token=t(uu["prompt_token"]),
                               prompt_speech_feat=t(uu["prompt_feat"]), source_speech_token=t(uu["token"]), stream=True):
                pass
            out[f"{n_tok}_{n_prompt}"] = np.array(calls, np.int64)
            print("tts schedule", n_tok, n_prompt, calls)
    finally:
        ref_model_mod.time.sleep = real_sleep
    np.savez_compressed(os.path.join(OUT, "tts_schedule.npz"), **out)


if __name__ == "__main__":
    main()
