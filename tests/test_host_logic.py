"""CPU: host-side logic of the engine (schedule, weight packing) against the oracle."""
import numpy as np
import torch

import token2wav_oracle as O


def test_euler_schedule_matches_reference_running_t():
    from cosyvoice2_eu_b200.engine import euler_schedule
    ts, dts = euler_schedule(10)
    span = O.t_schedule(10)
    t, dt = span[0:1], span[1] - span[0]
    for step in range(1, 11):
        assert ts[step - 1] == np.float32(t.item())
        assert dts[step - 1] == np.float32(dt.item())
        t = t + dt
        if step < 10:
            dt = span[step + 1] - t


def test_pack_conv_transpose_equals_conv_transpose1d():
    """Phase-concatenated GEMM weights reproduce F.conv_transpose1d (k/stride of the three HiFT upsamplers)."""
    from cosyvoice2_eu_b200 import pack
    g = torch.Generator().manual_seed(0)
    for (k, s, p) in ((16, 8, 4), (11, 5, 3), (7, 3, 2)):
        cin, cout, T = 8, 4, 9
        wt = torch.randn(cin, cout, k, generator=g)
        x = torch.randn(1, cin, T, generator=g)
        ref = torch.nn.functional.conv_transpose1d(x, wt, stride=s, padding=p)[0]          # [cout, s*T]
        W = pack._convT_w(wt, s).float()                                                   # [s*cout, q*cin]
        q = W.shape[1] // cin
        xt = x[0].t()                                                                      # [T, cin]
        rows = T + q
        out = torch.zeros(rows * s - p + 64, cout)
        flat = torch.zeros((rows + 1) * s * cout + 4096)
        for m in range(rows):
            a = torch.cat([xt[m - j] if 0 <= m - j < T else torch.zeros(cin) for j in range(q)])
            y = W @ a                                                                      # [s*cout]
            e0 = m * s * cout - p * cout
            for n in range(s * cout):
                e = e0 + n
                if 0 <= e < T * s * cout:
                    flat[e] = y[n]
        got = flat[:T * s * cout].reshape(T * s, cout).t()
        assert torch.allclose(got, ref.to(torch.float16).float(), atol=2e-2), (k, s, p)


def test_pack_names_cover_state_dict(fixture_weights):
    from cosyvoice2_eu_b200 import pack
    fs, hs = fixture_weights
    pf, ph = pack.pack_flow(fs), pack.pack_hift(hs)
    assert pf["est.res.0.c1.w"].shape == (256, 3 * 320) and pf["est.res.13.c1.w"].shape == (256, 3 * 512)
    assert pf["est.tfm.5.2.qkv.w"].shape == (1536, 256) and pf["enc.layers.3.qkv.w"].shape == (2048, 512)   # (q+u) | (q+v) | k | v
    assert ph["hift.ups.0.w"].shape == (8 * 256, 2 * 512) and ph["hift.ups.1.w"].shape == (5 * 128, 3 * 256)
    assert ph["hift.ups.2.w"].shape == (3 * 64, 3 * 128) and ph["hift.conv_pre.w"].shape == (512, 7 * 128)
    assert ph["f0.c0.w"].shape == (3, 80, 512) and ph["hift.sd.0.w"].shape == (30, 18, 256)
    # weight-norm fold agrees with the oracle's
    w = O.fold_weight_norm(hs, "conv_post")
    assert torch.allclose(ph["hift.conv_post.w"].float().reshape(18, 7, 64).permute(0, 2, 1), w, atol=2e-3)


# ------------------------------------------------------------------------------------------ StreamScheduler
class _StubT2W:
    """Records what the scheduler asks for; stands in for B200Token2Wav (no GPU)."""

    def __init__(self):
        import types
        self.flow = types.SimpleNamespace(pre_lookahead_len=3)
        self.hift_cache_dict = {}
        self.calls = []

    def token2wav_stream_batch(self, requests, finalize, noises=None):
        import torch
        self.calls.append((finalize, [(r["uuid"], int(r["token"].shape[1]), int(r["token_offset"])) for r in requests]))
        for r in requests:
            assert r["token"].dtype == torch.int32 and r["token"].dim() == 2
        return [torch.full((1, 4), float(r["token"].shape[1])) for r in requests]


def _per_session(calls):
    seen = {}
    for fin, reqs in calls:
        for uuid, n_vis, off in reqs:
            seen.setdefault(uuid, []).append((n_vis, off, fin))
    return seen


def test_stream_scheduler_reproduces_the_reference_chunk_schedule():
    """Whatever the arrival pattern of the tokens, every session gets exactly the calls CosyVoice2Model.tts(stream=True) would make
    (CV/cli/model.py:351-381, restated as token2wav_oracle.stream_schedule), and sessions that are ready together share a call."""
    import numpy as np
    import torch
    import token2wav_oracle as O
    from cosyvoice2_eu_b200 import StreamScheduler
    specs = {"a": (70, 10), "b": (95, 25), "c": (55, 12), "d": (5, 0), "e": (28, 50)}
    rng = np.random.Generator(np.random.Philox(key=5))
    t2w = _StubT2W()
    sch = StreamScheduler(t2w, token_hop_len=25)
    left = {}
    for u, (n, p) in specs.items():
        sch.open(u, torch.zeros(1, p, dtype=torch.int32), torch.zeros(1, 2 * p, 80), torch.zeros(1, 192))
        left[u] = list(range(n))
        assert t2w.hift_cache_dict[u] is None
    delivered = []
    ended_at_first = {}
    while sch.sessions:
        for u in list(left):
            k = int(rng.integers(0, 40))
            if left[u]:
                sch.push(u, left[u][:k])
                left[u] = left[u][k:]
            if not left[u]:
                sch.close(u)
                del left[u]
        while sch.pending():
            for u, s in sch.sessions.items():     # was the producer done when the first chunk went out? (model.py:369, stale hop)
                if s.token_offset == 0 and sch._ready(s):
                    ended_at_first[u] = s.ended
            delivered.extend(sch.step())
    got = _per_session(t2w.calls)
    for u, (n, p) in specs.items():
        assert got[u] == O.stream_schedule(n, p, all_tokens_ready=ended_at_first.get(u, False)), u
        assert u not in t2w.hift_cache_dict                          # dropped after the final chunk (model.py:395-396)
    assert [d[2] for d in delivered if d[0] == "a"] == [False] * (len(got["a"]) - 1) + [True]
    assert any(len(reqs) > 1 for _, reqs in t2w.calls)                # concurrent sessions were batched
    for fin, reqs in t2w.calls:
        assert len({r[0] for r in reqs}) == len(reqs)                # a session appears at most once per call


def test_schedules_match_the_references_own_tts_loop(golden):
    """tests/golden/tts_schedule.npz = the token2wav calls made by the reference's CosyVoice2Model.tts(stream=True) itself with all
    tokens present from the start (vc_job): the oracle restatement, the host helper and the StreamScheduler fed the same way
    must make exactly those calls (including the early finalization caused by the stale hop in the break test, model.py:369)."""
    import torch
    from cosyvoice2_eu_b200 import StreamScheduler
    from cosyvoice2_eu_b200.scheduler import chunk_schedule
    g = golden("tts_schedule")
    assert len(g) >= 10
    for key, ref in g.items():
        n, p = map(int, key.split("_"))
        want = [(int(a), int(b), bool(c)) for a, b, c in ref.tolist()]
        assert O.stream_schedule(n, p, all_tokens_ready=True) == want, key
        assert chunk_schedule(n, p, all_tokens_ready=True) == want, key
        assert chunk_schedule(n, p) == O.stream_schedule(n, p), key
        t2w = _StubT2W()
        sch = StreamScheduler(t2w, token_hop_len=25)
        sch.open("u", torch.zeros(1, p, dtype=torch.int32), torch.zeros(1, 2 * p, 80), torch.zeros(1, 192))
        sch.push("u", list(range(n)))
        sch.close("u")
        while sch.pending():
            sch.step()
        assert _per_session(t2w.calls)["u"] == want, key
    # the (70, 10) session differs between the two arrival patterns: that is the case the stale hop decides
    assert O.stream_schedule(70, 10) != O.stream_schedule(70, 10, all_tokens_ready=True)


def test_stream_scheduler_is_event_driven_across_threads():
    """Producer threads push tokens one by one; the consumer sleeps on the condition variable (no polling) and still makes the
    reference's calls for every session."""
    import threading
    import time
    import torch
    import token2wav_oracle as O
    from cosyvoice2_eu_b200 import StreamScheduler
    specs = {"x": (64, 10), "y": (31, 0), "z": (90, 25)}
    t2w = _StubT2W()
    sch = StreamScheduler(t2w, token_hop_len=25)
    for u, (n, p) in specs.items():
        sch.open(u, torch.zeros(1, p, dtype=torch.int32), torch.zeros(1, 2 * p, 80), torch.zeros(1, 192))

    def produce(u, n):
        for i in range(n):
            sch.push(u, i)
            if i % 16 == 0:
                time.sleep(0.001)
        sch.close(u)

    threads = [threading.Thread(target=produce, args=(u, n)) for u, (n, _) in specs.items()]
    chunks = []
    for t in threads:
        t.start()
    sch.run(lambda u, sp, fin: chunks.append((u, fin)), idle_timeout=5.0)
    for t in threads:
        t.join()
    assert not sch.sessions
    got = _per_session(t2w.calls)
    for u, (n, p) in specs.items():
        assert got[u] == O.stream_schedule(n, p), u
    assert sorted(u for u, fin in chunks if fin) == sorted(specs)
