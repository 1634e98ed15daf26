"""CPU: the oracle restatement (oracle/token2wav_oracle.py) against golden vectors that
oracle/make_golden.py produced from the UNMODIFIED reference.  This pins the oracle."""
import numpy as np
import torch

import token2wav_oracle as O
import weights


def snr_db(ref, x):
    ref, x = np.asarray(ref, np.float64), np.asarray(x, np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-30))


def T(x):
    return torch.from_numpy(x)


def test_masks_bit_exact(golden):
    g = golden("masks")
    lens = T(g["lens"])
    assert np.array_equal(O.pad_mask(lens).numpy(), g["pad_mask"])
    assert np.array_equal(O.chunk_mask(13, 5).numpy(), g["chunk_mask_13_5"])
    valid = ~O.pad_mask(lens, 12)[:, None, :]
    assert np.array_equal(O.attention_mask(valid, 5).numpy(), g["att_mask_chunk5"])
    assert np.array_equal(O.attention_mask(valid, 0).numpy(), g["att_mask_full"])


def test_stream_schedule_matches_reference_trace(golden):
    # SURVEY.md 3.2 probe trace for N=200, P=75: prefixes 28..178 then 200
    s = O.stream_schedule(200, 75)
    assert [a for a, _, _ in s] == [28, 53, 78, 103, 128, 153, 178, 200]
    assert [b for _, b, _ in s] == [0, 25, 50, 75, 100, 125, 150, 175]
    g = golden("stream")
    assert np.array_equal(np.array([(a, b, int(c)) for a, b, c in O.stream_schedule(70, 10)]), g["schedule"])


def test_estimator_single_call(golden, fixture_weights):
    g = golden("est")
    sd = O._sub(fixture_weights[0], "decoder.estimator.")
    with torch.inference_mode():
        for streaming, key in ((False, "out_offline"), (True, "out_streaming")):
            out = O.estimator_forward(sd, T(g["x"]), T(g["mask"]), T(g["mu"]), T(g["t"]), T(g["spks"]), T(g["cond"]),
                                      streaming=streaming)
            assert np.abs(out.numpy() - g[key]).max() < 2e-5


def test_tiny_flow_hift_token2wav(golden, fixture_weights):
    g = golden("tiny")
    fs, hs = fixture_weights
    u = weights.make_utterance(int(g["n_tok"]), int(g["n_prompt"]), int(g["seed"]))
    with torch.inference_mode():
        mel, inter = O.flow_inference(fs, weights.cfm_rand_noise(), T(u["token"]), T(u["prompt_token"]), T(u["prompt_feat"]),
                                      T(u["embedding"]), return_intermediates=True)
        assert np.abs(mel.numpy() - g["mel"]).max() < 1e-4
        assert np.abs(inter["h"].numpy() - g["encoder_out"]).max() < 1e-4
        mel_s = O.flow_inference(fs, weights.cfm_rand_noise(), T(u["token"]), T(u["prompt_token"]), T(u["prompt_feat"]),
                                 T(u["embedding"]), streaming=True, finalize=False)
        assert np.abs(mel_s.numpy() - g["mel_stream_nonfinal"]).max() < 1e-4
        # hift fed the REFERENCE mel + identical injected noise (stage-wise gate)
        noise = T(weights.make_nsf_noise(g["mel"].shape[2] * 480, int(g["seed"])))
        wav, src, hi = O.hift_inference(hs, T(g["mel"]), None, noise, return_intermediates=True)
        assert np.abs(hi["f0"].numpy() - g["f0"]).max() < 1e-2
        assert np.abs(O.source_stft(T(g["source"]).squeeze(1)).numpy() - g["s_stft"]).max() < 1e-5
        print("f0 maxdiff", np.abs(hi["f0"].numpy() - g["f0"]).max(), "source snr", snr_db(g["source"], src.numpy()),
              "wav snr", snr_db(g["wav"], wav.numpy()))
        assert snr_db(g["source"], src.numpy()) > 60
        assert snr_db(g["wav"], wav.numpy()) > 50


def test_streaming_token2wav(golden, fixture_weights):
    g = golden("stream")
    fs, hs = fixture_weights
    seed = int(g["seed"])
    u = weights.make_utterance(int(g["n_tok"]), int(g["n_prompt"]), seed)
    eng = O.OracleToken2Wav(fs, hs, weights.cfm_rand_noise())
    eng.hift_cache_dict["s"] = None
    for ci, (n_vis, off, fin) in enumerate(g["schedule"]):
        noise = T(weights.make_nsf_noise(int(g["mel_lens"][ci]) * 480, seed * 100 + ci))
        w = eng.token2wav(T(u["token"][:, :n_vis]), T(u["prompt_token"]), T(u["prompt_feat"]), T(u["embedding"]),
                          token_offset=int(off), uuid="s", stream=not bool(fin), finalize=bool(fin), noise=noise)
        ref = g[f"chunk{ci}"]
        assert w.shape == ref.shape            # shapes bit-exact
        assert snr_db(ref, w.numpy()) > 40


def test_streaming_stages_pin_the_oracle(golden, fixture_weights):
    """Stage-wise pin of the oracle's streaming pieces against the reference's own intermediate values
    (tests/golden/stream_stages.npz): per-chunk flow mel, hift with the cache_source overwrite, fade_in_out."""
    g = golden("stream_stages")
    fs, hs = fixture_weights
    seed = int(g["seed"])
    u = weights.make_utterance(int(g["n_tok"]), int(g["n_prompt"]), seed)
    window = np.hamming(2 * 3840)
    with torch.inference_mode():
        for ci, (n_vis, off, fin) in enumerate(g["schedule"]):
            mel = O.flow_inference(fs, weights.cfm_rand_noise(), T(u["token"][:, :n_vis]), T(u["prompt_token"]), T(u["prompt_feat"]),
                                   T(u["embedding"]), streaming=not bool(fin), finalize=bool(fin))
            assert np.abs(mel.numpy() - g[f"flow_mel{ci}"]).max() < 1e-4
            hm, cs = T(g[f"hift_mel{ci}"]), T(g[f"cache_source{ci}"])
            noise = T(weights.make_nsf_noise(hm.shape[2] * 480, seed * 100 + ci))
            speech, source = O.hift_inference(hs, hm, cs, noise)
            assert snr_db(g[f"speech_pre{ci}"], speech.numpy()) > 50
            assert snr_db(g[f"source{ci}"], source.numpy()) > 60
            if ci > 0:
                assert torch.equal(source[:, :, :3840], cs)
                post = O.fade_in_out(T(g[f"speech_pre{ci}"]), T(g[f"fade_old{ci}"]), window)
                assert np.array_equal(post.numpy(), g[f"speech_post{ci}"])          # same float64 window maths: bit-exact


def test_long_hift_pins_the_oracle(golden, fixture_weights):
    """configs[1] length (500 mel frames): oracle vocoder on the reference mel vs the reference waveform (the 650 / 1150-frame
    flow passes are compared on the GPU side only -- they cost the CPU suite half a minute each)."""
    g = golden("long")
    hs = fixture_weights[1]
    noise = T(weights.make_nsf_noise(g["cfg2_mel"].shape[2] * 480, int(g["cfg2_seed"])))
    with torch.inference_mode():
        wav, _, hi = O.hift_inference(hs, T(g["cfg2_mel"]), None, noise, return_intermediates=True)
    assert np.abs(hi["f0"].numpy() - g["cfg2_f0"]).max() < 1e-2
    assert snr_db(g["cfg2_wav"], wav.numpy()) > 50
