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
