"""Generate tests/golden/*.npz from the UNMODIFIED reference (container only).

    python oracle/make_golden.py            # needs /root/reference

Every tensor stored here is an output of the reference's own modules
(CausalMaskedDiffWithXvec / HiFTGenerator / CosyVoice2Model.token2wav, imported
read-only through oracle/ref_shims.py) with the deterministic fixture weights of
oracle/weights.py loaded via load_state_dict, on the synthetic inputs of
oracle/weights.make_utterance and with the NSF Gaussian noise of
oracle/weights.make_nsf_noise injected in place of torch.randn_like
(CV/hifigan/generator.py:334).  The oracle restatement and the CUDA engine are
both checked against these files.

Cases
  tiny   : P=10 prompt tokens, N=30 tokens (T=80 mel frames, 60 generated, 28 800 samples)
  cfg1   : BASELINE.json configs[0]: P=75, N=200 (T=550, 400 generated, 192 000 samples = 8 s)
  est    : one estimator call in the shape of CV/bin/export_onnx.py:35-42 (uniform inputs, batch 2)
  stream : P=10, N=70, hop 25 -> the chunk schedule of CV/cli/model.py:351-381 driven through token2wav
  masks  : integer/bool bookkeeping of CV/utils/mask.py
"""
import contextlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import weights  # noqa: E402
from token2wav_oracle import stream_schedule  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


@contextlib.contextmanager
def inject_noise(noises):
    """Replace torch.randn_like for [B, L, 9] tensors by the next injected noise."""
    it = iter(noises)
    orig = torch.randn_like

    def patched(x, *a, **k):
        if x.dim() == 3 and x.shape[-1] == 9:
            n = next(it)
            assert tuple(n.shape) == tuple(x.shape), (n.shape, x.shape)
            return n.to(x)
        return orig(x, *a, **k)

    torch.randn_like = patched
    try:
        yield
    finally:
        torch.randn_like = orig


def t(x):
    return torch.from_numpy(x)


def flow_call(flow, u, streaming=False, finalize=True):
    n, p = u["token"].shape[1], u["prompt_token"].shape[1]
    mel, _ = flow.inference(token=t(u["token"]), token_len=torch.tensor([n], dtype=torch.int32),
                            prompt_token=t(u["prompt_token"]), prompt_token_len=torch.tensor([p], dtype=torch.int32),
                            prompt_feat=t(u["prompt_feat"]), prompt_feat_len=torch.tensor([2 * p], dtype=torch.int32),
                            embedding=t(u["embedding"]), streaming=streaming, finalize=finalize)
    return mel


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    flow, hift = ref_shims.build_reference_modules()
    flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
    hift.load_state_dict(weights.to_torch(weights.make_hift_state()))
    model = ref_shims.build_reference_model(flow, hift)
    model.device = torch.device("cpu")

    # ---- tiny + cfg1: flow, hift, token2wav offline ---------------------------------------
    for name, (n_tok, n_prompt, seed) in {"tiny": (30, 10, 1), "cfg1": (200, 75, 2)}.items():
        u = weights.make_utterance(n_tok, n_prompt, seed)
        with torch.inference_mode():
            mel = flow_call(flow, u)
            mel_stream = flow_call(flow, u, streaming=True, finalize=False) if name == "tiny" else None
            noise = t(weights.make_nsf_noise(mel.shape[2] * 480, seed))
            f0 = hift.f0_predictor(mel)
            with inject_noise([noise]):
                wav, source = hift.inference(speech_feat=mel)
            model.hift_cache_dict["g"] = None
            with inject_noise([noise]):
                wav_t2w = model.token2wav(t(u["token"]), t(u["prompt_token"]), t(u["prompt_feat"]), t(u["embedding"]),
                                          token_offset=0, uuid="g", finalize=True)
        assert torch.equal(wav, wav_t2w)
        d = dict(n_tok=n_tok, n_prompt=n_prompt, seed=seed, mel=mel.numpy(), f0=f0.numpy(), wav=wav.numpy())
        if name == "tiny":
            d.update(source=source.numpy(), mel_stream_nonfinal=mel_stream.numpy())
            # encoder / conditioning intermediates for module-level parity
            from cosyvoice.utils.mask import make_pad_mask
            with torch.inference_mode():
                tok = torch.cat([t(u["prompt_token"]), t(u["token"])], 1)
                emb = flow.input_embedding(torch.clamp(tok, min=0).long())
                h, _ = flow.encoder(emb, torch.tensor([tok.shape[1]]), streaming=False)
                d["encoder_out"] = h.numpy()
                # hift sub-modules
                s_stft = torch.cat(hift._stft(source.squeeze(1)), dim=1)
                d["s_stft"] = s_stft.numpy()
        print(name, "mel", mel.shape, "std %.3f" % mel.std(), "f0 %.1f..%.1f voiced %.2f" % (f0.min(), f0.max(), (f0 > 10).float().mean()),
              "wav absmax %.3f clamp %.4f" % (wav.abs().max(), (wav.abs() >= 0.99).float().mean()))
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **d)

    # ---- estimator single call (export_onnx.py:35-42 shape) --------------------------------
    g = np.random.Generator(np.random.Philox(key=77))
    T = 96
    x = g.random((2, 80, T), dtype=np.float32)
    mu = g.random((2, 80, T), dtype=np.float32)
    cond = g.random((2, 80, T), dtype=np.float32)
    spks = g.random((2, 80), dtype=np.float32)
    tt = g.random((2,), dtype=np.float32)
    mask = np.ones((2, 1, T), np.float32)
    with torch.inference_mode():
        out_off = flow.decoder.estimator(t(x), t(mask), t(mu), t(tt), t(spks), t(cond), streaming=False)
        out_str = flow.decoder.estimator(t(x), t(mask), t(mu), t(tt), t(spks), t(cond), streaming=True)
    np.savez_compressed(os.path.join(OUT, "est.npz"), x=x, mu=mu, cond=cond, spks=spks, t=tt, mask=mask,
                        out_offline=out_off.numpy(), out_streaming=out_str.numpy())
    print("est", out_off.shape, "std %.3f" % out_off.std())

    # ---- streaming schedule through token2wav ------------------------------------------------
    n_tok, n_prompt, seed = 70, 10, 3
    u = weights.make_utterance(n_tok, n_prompt, seed)
    sched = stream_schedule(n_tok, n_prompt)
    model.hift_cache_dict["s"] = None
    chunks, mel_lens = [], []
    with torch.inference_mode():
        for ci, (n_vis, off, fin) in enumerate(sched):
            # mel handed to hift = new frames (+ 8 cached) ; its length fixes the noise shape
            n_new = (n_vis - (0 if fin else 3)) * 2 - off * 2
            mel_len = n_new + (8 if model.hift_cache_dict["s"] is not None else 0)
            mel_lens.append(mel_len)
            noise = t(weights.make_nsf_noise(mel_len * 480, seed * 100 + ci))
            with inject_noise([noise]):
                w = model.token2wav(t(u["token"][:, :n_vis]), t(u["prompt_token"]), t(u["prompt_feat"]), t(u["embedding"]),
                                    token_offset=off, uuid="s", stream=not fin, finalize=fin)
            chunks.append(w.numpy())
    d = dict(n_tok=n_tok, n_prompt=n_prompt, seed=seed, schedule=np.array([(a, b, int(c)) for a, b, c in sched], np.int64),
             mel_lens=np.array(mel_lens, np.int64))
    for i, c in enumerate(chunks):
        d[f"chunk{i}"] = c
    print("stream", sched, [c.shape[1] for c in chunks], "total", sum(c.shape[1] for c in chunks))
    np.savez_compressed(os.path.join(OUT, "stream.npz"), **d)

    # ---- integer bookkeeping -------------------------------------------------------------------
    from cosyvoice.utils.mask import add_optional_chunk_mask, make_pad_mask, subsequent_chunk_mask
    lens = torch.tensor([7, 3, 12, 1])
    pm = make_pad_mask(lens)
    cm = subsequent_chunk_mask(13, 5)
    valid = (~make_pad_mask(lens, 12)).unsqueeze(1)
    aocm = add_optional_chunk_mask(torch.zeros(4, 12, 1), valid, False, False, 0, 5, -1)
    aocm0 = add_optional_chunk_mask(torch.zeros(4, 12, 1), valid, False, False, 0, 0, -1)
    np.savez_compressed(os.path.join(OUT, "masks.npz"), lens=lens.numpy(), pad_mask=pm.numpy(), chunk_mask_13_5=cm.numpy(),
                        att_mask_chunk5=aocm.numpy(), att_mask_full=aocm0.numpy())
    print("masks ok")


if __name__ == "__main__":
    main()
