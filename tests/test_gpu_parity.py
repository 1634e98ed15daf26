"""GPU parity of the engine against golden vectors produced by the UNMODIFIED reference (tests/golden/*.npz,
oracle/make_golden.py) and against the oracle restatement, through the reference-shaped Python host which calls
the C ABI.  Tolerances are north_star's: mel max-abs <= 1e-2, waveform SNR >= 35 dB (stage-wise, SURVEY.md 7.3:
hift is fed the reference mel + the identical injected NSF noise), integer bookkeeping bit-exact."""
import numpy as np
import pytest
import torch

import weights

pytestmark = pytest.mark.gpu

MEL_TOL = 1e-2
SNR_MIN = 35.0


def snr_db(ref, x):
    ref, x = np.asarray(ref, np.float64), np.asarray(x, np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-30))


def T(x):
    return torch.from_numpy(np.asarray(x))


@pytest.fixture(scope="module")
def engine(fixture_weights):
    from cosyvoice2_eu_b200 import B200Flow, B200HiFT, B200Token2Wav
    fs, hs = fixture_weights
    flow, hift = B200Flow("cuda:0"), B200HiFT("cuda:0")
    flow.load_state_dict(fs)
    hift.load_state_dict(hs)
    return flow, hift, B200Token2Wav(flow, hift)


def test_estimator_single_call(engine, golden):
    """The reference's only numeric check on this path has this shape (cosyvoice/bin/export_onnx.py:117-133)."""
    flow = engine[0]
    g = golden("est")
    for streaming, key in ((False, "out_offline"), (True, "out_streaming")):
        out = flow.estimator_forward(T(g["x"]), T(g["mask"]), T(g["mu"]), T(g["t"]), T(g["spks"]), T(g["cond"]), streaming=streaming)
        err = np.abs(out.cpu().numpy() - g[key]).max()
        print("estimator", key, "max-abs err", err, "ref std", g[key].std())
        assert err < 5e-3


def test_estimator_ragged_mask(engine, golden, fixture_weights):
    """Padded batch rows must reproduce the unpadded B=1 result (length-aware kernels)."""
    import token2wav_oracle as O
    flow = engine[0]
    g = golden("est")
    n = 57
    mask = g["mask"].copy()
    mask[1, :, n:] = 0
    out = flow.estimator_forward(T(g["x"]), T(mask), T(g["mu"]), T(g["t"]), T(g["spks"]), T(g["cond"])).cpu()
    sd = O._sub(fixture_weights[0], "decoder.estimator.")
    with torch.inference_mode():
        ref1 = O.estimator_forward(sd, T(g["x"][1:2, :, :n]), T(mask[1:2, :, :n]), T(g["mu"][1:2, :, :n]), T(g["t"][1:2]),
                                   T(g["spks"][1:2]), T(g["cond"][1:2, :, :n]))
    assert float((out[1:2, :, :n] - ref1).abs().max()) < 5e-3
    assert float(out[1, :, n:].abs().max()) == 0.0
    assert np.abs(out[0].numpy() - g["out_offline"][0]).max() < 5e-3


def _utt(g):
    return {k: T(v) for k, v in weights.make_utterance(int(g["n_tok"]), int(g["n_prompt"]), int(g["seed"])).items()}


def test_flow_tiny(engine, golden):
    flow = engine[0]
    g = golden("tiny")
    u = _utt(g)
    (mel, inter, lens) = flow.inference_batch([u["token"][0]], [u["prompt_token"][0]], [u["prompt_feat"][0]], [u["embedding"][0]],
                                              return_intermediates=True)
    enc_err = np.abs(inter["encoder_out"].cpu().numpy() - g["encoder_out"]).max()
    err = np.abs(mel.cpu().numpy() - g["mel"]).max()
    print("tiny: encoder_out max-abs err", enc_err, "mel max-abs err", err)
    assert mel.shape == g["mel"].shape
    assert enc_err < 2e-2
    assert err <= MEL_TOL
    mel_s, _ = flow.inference(u["token"], None, u["prompt_token"], None, u["prompt_feat"], None, u["embedding"], True, False)
    err_s = np.abs(mel_s.cpu().numpy() - g["mel_stream_nonfinal"]).max()
    print("tiny streaming non-final mel err", err_s)
    assert mel_s.shape == g["mel_stream_nonfinal"].shape
    assert err_s <= MEL_TOL


def test_hift_tiny(engine, golden):
    hift = engine[1]
    g = golden("tiny")
    noise = T(weights.make_nsf_noise(g["mel"].shape[2] * 480, int(g["seed"])))
    speech, source, f0 = hift.inference(T(g["mel"]), noise=noise, return_f0=True)
    f0_err = np.abs(f0.cpu().numpy() - g["f0"]).max()
    s_src = snr_db(g["source"], source.cpu().numpy())
    s_wav = snr_db(g["wav"], speech.cpu().numpy())
    print("tiny hift: f0 max err", f0_err, "source SNR", s_src, "wav SNR", s_wav)
    assert speech.shape == g["wav"].shape
    assert f0_err < 5e-2
    assert s_src >= 45.0
    assert s_wav >= SNR_MIN


def test_token2wav_tiny_end_to_end(engine, golden):
    t2w = engine[2]
    g = golden("tiny")
    u = _utt(g)
    noise = T(weights.make_nsf_noise(g["mel"].shape[2] * 480, int(g["seed"])))
    t2w.hift_cache_dict["t"] = None
    wav = t2w.token2wav(u["token"], u["prompt_token"], u["prompt_feat"], u["embedding"], 0, "t", finalize=True, noise=noise)
    assert tuple(wav.shape) == g["wav"].shape
    # end-to-end SNR is governed by F0->phase drift of the mel error; reported, loosely gated
    s = snr_db(g["wav"], wav.cpu().numpy())
    print("tiny token2wav end-to-end SNR", s)
    assert s > 5.0


def test_cfg1_full_size(engine, golden):
    """BASELINE.json configs[0]: 200 tokens + 75-token prompt -> 8 s."""
    flow, hift, _ = engine
    g = golden("cfg1")
    u = _utt(g)
    mel, _ = flow.inference(u["token"], None, u["prompt_token"], None, u["prompt_feat"], None, u["embedding"], False, True)
    err = np.abs(mel.cpu().numpy() - g["mel"]).max()
    noise = T(weights.make_nsf_noise(g["mel"].shape[2] * 480, int(g["seed"])))
    speech, _, f0 = hift.inference(T(g["mel"]), noise=noise, return_f0=True)
    s_wav = snr_db(g["wav"], speech.cpu().numpy())
    print("cfg1: mel max-abs err", err, "f0 err", np.abs(f0.cpu().numpy() - g["f0"]).max(), "hift SNR", s_wav)
    assert err <= MEL_TOL
    assert s_wav >= SNR_MIN


def test_batch_equals_single(engine):
    """Batched ragged inference == per-utterance B=1 inference (the parity target)."""
    flow, hift, t2w = engine
    utts = [_utt(dict(n_tok=n, n_prompt=p, seed=s)) for n, p, s in ((30, 10, 1), (57, 20, 4), (41, 7, 5))]
    mel_b, lens = flow.inference_batch([u["token"][0] for u in utts], [u["prompt_token"][0] for u in utts],
                                       [u["prompt_feat"][0] for u in utts], [u["embedding"][0] for u in utts])
    assert lens.tolist() == [60, 114, 82]
    noises = [T(weights.make_nsf_noise(int(n) * 480, 9 + i)) for i, n in enumerate(lens)]
    nz = torch.zeros(3, 480 * mel_b.shape[2], 9)
    for i, n in enumerate(noises):
        nz[i, :n.shape[1]] = n[0]
    wav_b, _ = hift.inference(mel_b, noise=nz, lens=lens)
    for i, u in enumerate(utts):
        mel_1, _ = flow.inference(u["token"], None, u["prompt_token"], None, u["prompt_feat"], None, u["embedding"], False, True)
        n = int(lens[i])
        d = float((mel_b[i, :, :n] - mel_1[0]).abs().max())
        assert float(mel_b[i, :, n:].abs().max()) == 0.0 if n < mel_b.shape[2] else True
        wav_1, _ = hift.inference(mel_1, noise=noises[i])
        dw = float((wav_b[i, :480 * n] - wav_1[0]).abs().max())
        print("utt", i, "batch-vs-single mel", d, "wav", dw)
        assert d < 1e-4
        assert snr_db(wav_1[0].cpu().numpy(), wav_b[i, :480 * n].cpu().numpy()) > 60


def test_streaming_schedule(engine, golden):
    """Chunk schedule of CosyVoice2Model.tts(stream=True) driven through token2wav: shapes bit-exact."""
    import token2wav_oracle as O
    t2w = engine[2]
    g = golden("stream")
    seed = int(g["seed"])
    u = _utt(g)
    sched = O.stream_schedule(int(g["n_tok"]), int(g["n_prompt"]))
    assert np.array_equal(np.array([(a, b, int(c)) for a, b, c in sched]), g["schedule"])
    t2w.hift_cache_dict["s"] = None
    for ci, (n_vis, off, fin) in enumerate(sched):
        noise = T(weights.make_nsf_noise(int(g["mel_lens"][ci]) * 480, seed * 100 + ci))
        w = t2w.token2wav(u["token"][:, :n_vis], u["prompt_token"], u["prompt_feat"], u["embedding"], token_offset=off, uuid="s",
                          stream=not fin, finalize=fin, noise=noise)
        ref = g[f"chunk{ci}"]
        assert tuple(w.shape) == ref.shape
        print("stream chunk", ci, "shape", tuple(w.shape), "e2e SNR", snr_db(ref, w.cpu().numpy()))


def test_cuda_graph_replay_matches_eager(engine):
    """One captured graph serves every utterance of its shape bucket (lengths are device side)."""
    from cosyvoice2_eu_b200 import GraphedToken2Wav
    flow, hift, t2w = engine
    g = GraphedToken2Wav(t2w)
    for n, p, s in ((30, 10, 1), (25, 12, 6), (31, 9, 7)):       # same (32, 32) bucket
        u = _utt(dict(n_tok=n, n_prompt=p, seed=s))
        wav, lens = g([u["token"][0]], [u["prompt_token"][0]], [u["prompt_feat"][0]], [u["embedding"][0]])
        torch.cuda.synchronize()
        mel_e, _ = flow.inference(u["token"], None, u["prompt_token"], None, u["prompt_feat"], None, u["embedding"], False, True)
        n_mel = 2 * n
        assert int(lens[0]) == 480 * n_mel
        assert float((g.last_mel[0, :, :n_mel] - mel_e[0]).abs().max()) < 1e-5
        assert bool(torch.isfinite(wav).all()) and float(wav[0, :480 * n_mel].abs().max()) > 1e-3
        assert float(wav[0, 480 * n_mel:].abs().max()) == 0.0
    assert len(g.graphs) == 1


def test_streaming_batched_sessions_equal_single_sessions(engine, golden):
    """configs[3]-style concurrency: several sessions stepped together through token2wav_stream_batch give exactly what each
    session gives alone through token2wav (and the golden chunk shapes for the (70, 10) session)."""
    import token2wav_oracle as O
    from cosyvoice2_eu_b200 import B200Token2Wav
    flow, hift, t2w = engine
    g = golden("stream")
    specs = [(70, 10, 3), (95, 25, 8), (55, 12, 9)]
    utts = [_utt(dict(n_tok=n, n_prompt=p, seed=s)) for n, p, s in specs]
    scheds = [O.stream_schedule(n, p) for n, p, _ in specs]
    noise_of = lambda si, ci, mel_len: T(weights.make_nsf_noise(mel_len * 480, 1000 * si + ci))
    # reference: each session alone
    solo = []
    for si, (u, sched) in enumerate(zip(utts, scheds)):
        one = B200Token2Wav(flow, hift)
        one.hift_cache_dict["x"] = None
        chunks = []
        for ci, (n_vis, off, fin) in enumerate(sched):
            n_new = (n_vis - (0 if fin else 3)) * 2 - off * 2
            mel_len = n_new + (8 if one.hift_cache_dict["x"] is not None else 0)
            chunks.append(one.token2wav(u["token"][:, :n_vis], u["prompt_token"], u["prompt_feat"], u["embedding"], off, "x",
                                        stream=not fin, finalize=fin, noise=noise_of(si, ci, mel_len)).cpu())
        solo.append(chunks)
    assert [tuple(c.shape) for c in solo[0]] == [g[f"chunk{i}"].shape for i in range(len(scheds[0]))]
    # batched: all sessions advance together
    multi = B200Token2Wav(flow, hift)
    for si in range(len(specs)):
        multi.hift_cache_dict[f"s{si}"] = None
    got = [[] for _ in specs]
    for step in range(max(len(s) for s in scheds)):
        for fin in (False, True):
            reqs, who, nz = [], [], []
            for si, (u, sched) in enumerate(zip(utts, scheds)):
                if step < len(sched) and sched[step][2] == fin:
                    n_vis, off, _ = sched[step]
                    n_new = (n_vis - (0 if fin else 3)) * 2 - off * 2
                    mel_len = n_new + (8 if multi.hift_cache_dict[f"s{si}"] is not None else 0)
                    reqs.append(dict(token=u["token"][:, :n_vis], prompt_token=u["prompt_token"], prompt_feat=u["prompt_feat"],
                                     embedding=u["embedding"], token_offset=off, uuid=f"s{si}"))
                    who.append(si)
                    nz.append(noise_of(si, step, mel_len))
            if reqs:
                outs = multi.token2wav_stream_batch(reqs, finalize=fin, noises=nz)
                for si, o in zip(who, outs):
                    got[si].append(o.cpu())
    for si in range(len(specs)):
        assert len(got[si]) == len(solo[si])
        for a, b in zip(got[si], solo[si]):
            assert a.shape == b.shape
            assert snr_db(b.numpy(), a.numpy()) > 60


def test_pcm16_epilogue_bit_exact(engine, golden):
    """int16 PCM written by the iSTFT kernel == the servers' `(speech * 2**15).astype(np.int16)`
    (runtime/python/fastapi/server.py:42) -- integer output, bit-exact."""
    hift = engine[1]
    g = golden("tiny")
    noise = T(weights.make_nsf_noise(g["mel"].shape[2] * 480, int(g["seed"])))
    speech, _, pcm = hift.inference(T(g["mel"]), noise=noise, return_pcm16=True)
    want = (speech.cpu().numpy() * (2 ** 15)).astype(np.int16)
    assert pcm.dtype == torch.int16 and tuple(pcm.shape) == tuple(speech.shape)
    assert np.array_equal(pcm.cpu().numpy(), want)
    assert int(np.abs(want).max()) > 30000          # the fixture reaches the +-0.99 clamp


def test_fused_euler_matches_separate_kernel(engine, golden, monkeypatch):
    """CFG combine + Euler update inside final_proj's epilogue (one K-concatenated GEMM over the cond/uncond pair)
    vs the separate euler_pack kernel: same mel within fp16-operand rounding of the projection."""
    import ctypes as C
    flow = engine[0]
    g = golden("tiny")
    u = _utt(g)
    args = ([u["token"][0]], [u["prompt_token"][0]], [u["prompt_feat"][0]], [u["embedding"][0]])
    mel_fused = flow.inference_batch(*args)[0].cpu().numpy()
    lib = flow.eng.lib
    lib.cv2_engine_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    assert lib.cv2_engine_set_option(flow.eng.h, b"fuse_euler", 0) == 0
    try:
        mel_sep = flow.inference_batch(*args)[0].cpu().numpy()
    finally:
        lib.cv2_engine_set_option(flow.eng.h, b"fuse_euler", 1)
    d = np.abs(mel_fused - mel_sep).max()
    print("fused vs separate Euler update: max-abs mel diff", d)
    assert d < 3e-3
    assert np.abs(mel_fused - g["mel"]).max() <= MEL_TOL and np.abs(mel_sep - g["mel"]).max() <= MEL_TOL


def test_cfg3_full_size_properties(engine):
    """BASELINE.json configs[2] at full size (64 utterances, U[4,20] s, length-sorted ragged batch): size-independent properties.
    (i) integer bookkeeping is exact (mel frames = 2 * tokens, samples = 480 * frames), (ii) padding of every output is exactly
    zero, (iii) the batched result of sampled utterances (shortest, median, longest) equals their B=1 result, (iv) the int16 PCM of
    the batch equals the NumPy conversion, (v) everything is finite and clamped to +-0.99."""
    flow, hift, t2w = engine
    rng = np.random.Generator(np.random.Philox(key=1000))
    n_tokens = sorted(int(round(25 * d)) for d in rng.uniform(4.0, 20.0, size=64))
    utts = [_utt(dict(n_tok=n, n_prompt=75, seed=100 + i)) for i, n in enumerate(n_tokens)]
    args = ([u["token"][0] for u in utts], [u["prompt_token"][0] for u in utts], [u["prompt_feat"][0] for u in utts],
            [u["embedding"][0] for u in utts])
    mel, mel_lens = flow.inference_batch(*args)
    assert mel_lens.tolist() == [2 * n for n in n_tokens]
    assert bool(torch.isfinite(mel).all())
    Tm = mel.shape[2]
    pick = [0, 31, 63]
    noise = torch.zeros(64, 480 * Tm, 9)
    for i in pick:
        nz = T(weights.make_nsf_noise(int(mel_lens[i]) * 480, 700 + i))
        noise[i, :nz.shape[1]] = nz[0]
    speech, source, pcm = hift.inference(mel, noise=noise, lens=mel_lens, return_pcm16=True)
    sp = speech.cpu().numpy()
    assert np.isfinite(sp).all() and np.abs(sp).max() <= 0.99 + 1e-7
    assert np.array_equal(pcm.cpu().numpy(), (sp * (2 ** 15)).astype(np.int16))
    for i in range(64):
        n = int(mel_lens[i])
        if n < Tm:
            assert float(mel[i, :, n:].abs().max()) == 0.0
            assert float(np.abs(sp[i, 480 * n:]).max()) == 0.0
    for i in pick:
        u = utts[i]
        n = int(mel_lens[i])
        mel_1, _ = flow.inference(u["token"], None, u["prompt_token"], None, u["prompt_feat"], None, u["embedding"], False, True)
        d = float((mel[i, :, :n] - mel_1[0]).abs().max())
        wav_1, _ = hift.inference(mel_1, noise=noise[i:i + 1, :480 * n])
        s = snr_db(wav_1[0].cpu().numpy(), sp[i, :480 * n])
        print("cfg3 utt", i, "tokens", n_tokens[i], "batch-vs-single mel", d, "wav SNR", s)
        assert tuple(mel_1.shape) == (1, 80, n)
        assert d < 1e-4
        assert s > 60


def test_2sm_kernels_match_1sm_kernels(engine, golden):
    """The 2-SM forms (tcgen05.mma.cta_group::2: fused FFN as CTA pairs, BN=256 GEMMs as CTA pairs) against the 1-SM kernels they
    replace on big launches: same contraction order per output, so the mel must agree to rounding noise; both within tolerance of
    the golden mel.  (Small launches always take the 1-SM kernels, so this runs a batch large enough to switch paths.)"""
    flow = engine[0]
    lib = flow.eng.lib
    utts = [_utt(dict(n_tok=240 + 3 * i, n_prompt=60, seed=40 + i)) for i in range(16)]    # 32 sequences x 5 tiles = 160 >= 148
    args = ([u["token"][0] for u in utts], [u["prompt_token"][0] for u in utts], [u["prompt_feat"][0] for u in utts],
            [u["embedding"][0] for u in utts])
    mel_2sm = flow.inference_batch(*args)[0].cpu().numpy()
    for name in (b"ffn_2cta", b"cluster_mc"):
        assert lib.cv2_engine_set_option(flow.eng.h, name, 0) == 0
    try:
        mel_1sm = flow.inference_batch(*args)[0].cpu().numpy()
    finally:
        for name in (b"ffn_2cta", b"cluster_mc"):
            lib.cv2_engine_set_option(flow.eng.h, name, 1)
    d = np.abs(mel_2sm - mel_1sm).max()
    print("2-SM vs 1-SM kernels: max-abs mel diff", d, "mel std", mel_1sm.std())
    assert np.isfinite(mel_2sm).all()
    assert d < 2e-3


@pytest.mark.parametrize("n_tok,n_prompt,seed", [(20, 0, 11), (1, 5, 12), (3, 0, 13), (2, 1, 14)])
def test_edge_cases_against_oracle(engine, fixture_weights, n_tok, n_prompt, seed):
    """No prompt at all (the SFT / instruct modes pass zeros(1, 0) tokens and zeros(1, 0, 80) mel, CV/cli/model.py:343-347 defaults),
    a single token, and the shortest mixes: flow mel and hift waveform against the oracle restatement."""
    import token2wav_oracle as O
    flow, hift, _ = engine
    fs, hs = fixture_weights
    u = _utt(dict(n_tok=n_tok, n_prompt=n_prompt, seed=seed))
    assert u["prompt_token"].shape == (1, n_prompt) and u["prompt_feat"].shape == (1, 2 * n_prompt, 80)
    mel, _ = flow.inference(u["token"], None, u["prompt_token"], None, u["prompt_feat"], None, u["embedding"], False, True)
    with torch.inference_mode():
        ref = O.flow_inference(fs, flow.rand_noise.cpu(), u["token"].long(), u["prompt_token"].long(), u["prompt_feat"], u["embedding"])
    assert tuple(mel.shape) == tuple(ref.shape) == (1, 80, 2 * n_tok)
    err = float((mel.cpu() - ref).abs().max())
    assert err < MEL_TOL, err
    noise = T(weights.make_nsf_noise(2 * n_tok * 480, seed))
    wav, _ = hift.inference(mel, noise=noise)
    with torch.inference_mode():
        wav_ref, _ = O.hift_inference(hs, mel.cpu(), noise=noise)
    assert tuple(wav.shape) == tuple(wav_ref.shape) == (1, 2 * n_tok * 480)
    s = snr_db(wav_ref.numpy(), wav.cpu().numpy())
    assert s >= SNR_MIN, s


def test_too_long_for_the_cfm_noise_buffer_is_rejected(engine):
    from cosyvoice2_eu_b200.lib import Cv2Error
    flow = engine[0]
    n = flow.rand_noise.shape[2] // 2 + 1
    with pytest.raises(Cv2Error):
        flow.inference_batch([torch.zeros(n, dtype=torch.int32)], [torch.zeros(0, dtype=torch.int32)], [torch.zeros(0, 80)],
                             [torch.zeros(192)])


@pytest.mark.parametrize("speed", [0.8, 1.25, 2.0])
def test_speed_change_matches_linear_interpolation(engine, golden, speed):
    """token2wav(speed != 1) (CV/cli/model.py:325-327): the mel is resampled in time with F.interpolate(mode='linear') before the
    vocoder.  Kernel vs the torch fp32 op on the same mel, and the whole call's output length."""
    import ctypes as C
    from cosyvoice2_eu_b200 import lib
    flow, hift, t2w = engine
    g = golden("tiny")
    mel = T(g["mel"]).cuda().contiguous()
    T_out = int(mel.shape[2] / speed)
    out = torch.empty(1, 80, T_out, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib.check(lib.load().cv2_mel_time_stretch(st, lib.ptr(mel), mel.shape[2], lib.ptr(out), T_out, 80))
    ref = torch.nn.functional.interpolate(mel, size=T_out, mode="linear")
    assert float((out - ref).abs().max()) < 1e-5
    u = _utt(g)
    wav = t2w.token2wav(u["token"], u["prompt_token"], u["prompt_feat"], u["embedding"], 0, "speed-test", stream=False, finalize=True,
                        speed=speed)
    assert wav.shape == (1, 480 * T_out)


def test_stream_scheduler_on_the_engine(engine, golden):
    """StreamScheduler (row F3) driving the real engine: three sessions fed token by token, chunks delivered through batched
    steps; the (70, 10) session's chunk shapes are the golden ones of the reference's own chunk loop, every session's audio adds
    up to 2 * 480 samples per token, and the vocoder caches are gone afterwards."""
    from cosyvoice2_eu_b200 import B200Token2Wav, StreamScheduler
    flow, hift, _ = engine
    g = golden("stream")
    t2w = B200Token2Wav(flow, hift)
    sch = StreamScheduler(t2w, token_hop_len=25)
    specs = {"s0": (70, 10, 3), "s1": (95, 25, 8), "s2": (55, 12, 9)}
    utts = {u: _utt(dict(n_tok=n, n_prompt=p, seed=s)) for u, (n, p, s) in specs.items()}
    for u, ut in utts.items():
        sch.open(u, ut["prompt_token"], ut["prompt_feat"], ut["embedding"])
    chunks = {u: [] for u in specs}
    pos = {u: 0 for u in specs}
    while sch.sessions:
        for u, (n, _, _) in specs.items():
            if pos[u] < n:
                k = min(7, n - pos[u])
                sch.push(u, utts[u]["token"][0, pos[u]:pos[u] + k].tolist())
                pos[u] += k
                if pos[u] == n:
                    sch.close(u)
        while sch.pending():
            for u, speech, fin in sch.step():
                chunks[u].append(speech.cpu())
    assert [tuple(c.shape) for c in chunks["s0"]] == [g[f"chunk{i}"].shape for i in range(len(chunks["s0"]))]
    for u, (n, _, _) in specs.items():
        assert sum(c.shape[1] for c in chunks[u]) == 2 * 480 * n
        # (a crossfaded head is w0 * new + w1 * old of two signals clamped to 0.99: Hamming halves sum to at most 1.08)
        assert all(bool(torch.isfinite(c).all()) and float(c.abs().max()) <= 0.99 * 1.09 for c in chunks[u])
        assert u not in t2w.hift_cache_dict
