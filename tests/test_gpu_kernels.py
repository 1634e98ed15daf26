"""GPU: single-kernel parity through the C ABI (cv2_op_*), each against a plain fp32 reference of the same op
(torch fp32 on the 16-bit-rounded operands for the tensor-core kernels; the oracle's functions for the rest)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from cosyvoice2_eu_b200 import lib
    return lib, lib.load()


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _act(x, act, f=0.0, a=None):
    F = torch.nn.functional
    if act == 1:
        return F.mish(x)
    if act == 2:
        return F.gelu(x)
    if act == 3:
        return F.silu(x)
    if act == 4:
        return F.elu(x)
    if act == 5:
        return F.leaky_relu(x, f)
    if act == 6:
        return x + torch.sin(x * a) ** 2 / (a + 1e-9)
    return x


def run_gemm(S, T_alloc, Kc, N, bn, taps, lens=None, ln=False, act=0, res=False, emit_ln=False, rowvec=False, mask_pre_res=False,
             accum=False, scale=1.0, seed=0):
    lib, L = _lib()
    g = torch.Generator().manual_seed(seed)
    dev = "cuda"
    kb = (Kc + 63) // 64
    A = (torch.randn(S, T_alloc, Kc, generator=g) * 0.5).half()
    W = torch.zeros(N, len(taps), kb * 64)
    W[:, :, :Kc] = torch.randn(N, len(taps), Kc, generator=g) / math.sqrt(Kc * len(taps))
    W = W.reshape(N, -1).half()
    bias = torch.randn(N, generator=g) * 0.1
    ln_g = 1 + 0.1 * torch.randn(N, generator=g)
    ln_b = 0.1 * torch.randn(N, generator=g)
    ln2_g = 1 + 0.1 * torch.randn(N, generator=g)
    ln2_b = 0.1 * torch.randn(N, generator=g)
    alpha = torch.exp(0.3 * torch.randn(N, generator=g))
    rv = torch.randn(S, N, generator=g) * 0.2
    R = torch.randn(S, T_alloc, N, generator=g)
    out_init = torch.randn(S, T_alloc, N, generator=g)
    lens_t = torch.tensor(lens if lens is not None else [T_alloc] * S, dtype=torch.int32)
    # ---- reference (fp32 on the rounded operands) ----
    Af = A.float()
    for s in range(S):
        Af[s, lens_t[s]:] = 0            # producers always write padded rows as zero
    A = Af.half()
    Wf = W.float().reshape(N, len(taps), kb * 64)[:, :, :Kc]
    acc = torch.zeros(S, T_alloc, N)
    for j, off in enumerate(taps):
        sh = torch.zeros_like(Af)
        if off >= 0:
            sh[:, :T_alloc - off] = Af[:, off:]
        else:
            sh[:, -off:] = Af[:, :T_alloc + off]
        acc += sh @ Wf[:, j].t()
    v = acc + bias
    if ln:
        v = torch.nn.functional.layer_norm(v, (N,), ln_g, ln_b, 1e-5)
    v = _act(v, act, 0.1, alpha)
    if rowvec:
        v = v + rv[:, None, :]
    valid = (torch.arange(T_alloc)[None, :] < lens_t[:, None])[:, :, None]
    if mask_pre_res:
        v = v * valid
    if res:
        v = v + R
    v = v * scale
    if accum:
        v = v + out_init
    ref32 = v
    ref16 = v * valid
    refln = torch.nn.functional.layer_norm(v, (N,), ln2_g, ln2_b, 1e-5) * valid
    # ---- device ----
    d = lambda t: t.to(dev).contiguous()
    A_d, W_d, bias_d = d(A), d(W), d(bias)
    out32 = d(out_init.clone()) if accum else torch.zeros(S, T_alloc, N, device=dev)
    out16 = torch.zeros(S, T_alloc, N, device=dev, dtype=torch.float16)
    outln = torch.zeros(S, T_alloc, N, device=dev, dtype=torch.float16)
    tens = dict(ln_g=d(ln_g), ln_b=d(ln_b), ln2_g=d(ln2_g), ln2_b=d(ln2_b), alpha=d(alpha), rv=d(rv), R=d(R), lens=d(lens_t))
    tap_arr = (C.c_int * len(taps))(*taps)
    p = lib.ptr
    lib.check(L.cv2_op_gemm_tap(_s(), p(A_d), S, T_alloc, Kc, Kc, p(W_d), N, W.shape[1], p(bias_d), bn, len(taps), tap_arr,
                                p(tens["lens"]), 0, p(tens["ln_g"]) if ln else None, p(tens["ln_b"]) if ln else None, 1e-5, act,
                                0.1, p(tens["alpha"]) if act == 6 else None, p(tens["rv"]) if rowvec else None, N,
                                int(mask_pre_res), p(tens["R"]) if res else None, scale, p(out32), int(accum), p(out16),
                                p(tens["ln2_g"]) if emit_ln else None, p(tens["ln2_b"]) if emit_ln else None,
                                p(outln) if emit_ln else None))
    torch.cuda.synchronize()
    # rows of active tiles only: t < len + 32 rounded up to the tile; compare valid rows for fp32 and all active rows for emits
    res_ = {}
    for s in range(S):
        n = int(lens_t[s])
        res_[s] = (float((out32[s, :n].cpu() - ref32[s, :n]).abs().max()),
                   float((out16[s, :n].float().cpu() - ref16[s, :n]).abs().max()),
                   float((outln[s, :n].float().cpu() - refln[s, :n]).abs().max()) if emit_ln else 0.0,
                   float(out16[s, n:min(T_alloc, n + 32)].float().abs().max()) if n < T_alloc else 0.0)
    return res_


@pytest.mark.parametrize("bn,N", [(256, 256), (128, 128), (64, 64), (128, 80), (64, 18), (256, 512)])
def test_gemm_plain(bn, N):
    r = run_gemm(2, 256, 256, N, bn, [0])
    for s, (e32, e16, _, pad) in r.items():
        assert e32 < 2e-3 and e16 < 5e-3 and pad == 0.0, r


def test_gemm_k_partial_and_taps():
    # K = 80 (second 64-block half out of bounds), 7 symmetric taps, ragged lengths
    r = run_gemm(3, 384, 80, 256, 256, [-3, -2, -1, 0, 1, 2, 3], lens=[384, 200, 77])
    for s, (e32, e16, _, pad) in r.items():
        assert e32 < 2e-3 and e16 < 5e-3 and pad == 0.0, r


def test_gemm_causal_ln_mish_rowvec_res_emitln():
    r = run_gemm(2, 256, 320, 256, 256, [-2, -1, 0], lens=[256, 130], ln=True, act=1, rowvec=True, mask_pre_res=True, res=True,
                 emit_ln=True)
    for s, (e32, e16, eln, pad) in r.items():
        assert e32 < 5e-3 and e16 < 1e-2 and eln < 2e-2 and pad == 0.0, r


@pytest.mark.parametrize("act", [2, 3, 4, 5, 6])
def test_gemm_activations(act):
    r = run_gemm(1, 128, 128, 128, 128, [-5, 0, 5], act=act)
    for s, (e32, e16, _, pad) in r.items():
        assert e32 < 3e-3 and e16 < 6e-3, r


def test_gemm_accumulate_scale_long_k():
    r = run_gemm(1, 256, 1024, 256, 256, [0], res=True, accum=True, scale=1.0 / 3.0)
    for s, (e32, e16, _, pad) in r.items():
        assert e32 < 3e-3, r


def _attn_ref(q, k, v, lens, chunk):
    S, H, T, D = q.shape
    out = torch.zeros(S, T, H * D)
    for s in range(S):
        n = int(lens[s])
        sc = q[s, :, :n].float() @ k[s, :, :n].float().transpose(-1, -2)
        if chunk > 0:
            i = torch.arange(n)
            vis = i[None, :] < ((i // chunk + 1) * chunk)[:, None]
            sc = sc.masked_fill(~vis[None], float("-inf"))
        o = torch.softmax(sc, -1) @ v[s, :, :n].float()
        out[s, :n] = o.transpose(0, 1).reshape(n, H * D)
    return out


@pytest.mark.parametrize("chunk", [0, 50])
def test_flash_attention(chunk):
    lib, L = _lib()
    g = torch.Generator().manual_seed(3)
    S, H, T, D = 3, 8, 384, 64
    lens = torch.tensor([384, 290, 100], dtype=torch.int32)
    q = (torch.randn(S, H, T, D, generator=g) * 0.125 * 2).half()
    k = (torch.randn(S, H, T, D, generator=g) * 2).half()
    v = torch.randn(S, H, T, D, generator=g).half()
    ref = _attn_ref(q, k, v, lens, chunk)
    qd, kd = q.cuda(), k.cuda()
    vtd = v.transpose(2, 3).contiguous().cuda()
    out = torch.zeros(S, T, H * D, dtype=torch.float16, device="cuda")
    lens_d = lens.cuda()
    lib.check(L.cv2_op_flash_attn(_s(), lib.ptr(qd), lib.ptr(kd), lib.ptr(vtd), lib.ptr(out), lib.ptr(lens_d), 0, S, H, T, chunk))
    torch.cuda.synchronize()
    err = float((out.float().cpu() - ref).abs().max())
    assert err < 5e-3, err


@pytest.mark.parametrize("chunk", [0, 50])
def test_flash_attention_long_ragged_and_lazy_rescale(chunk):
    """The estimator attention at the bench's tile count (T_alloc = 1152: 9 query tiles, 18 key tiles) with edge lengths (1, one key
    tile +- 1, a full allocation) and with logits that GROW along the keys, so that the running reference maximum has to be raised
    -- the lazy rescale (threshold 2^13, attention.cu) and its O / l rescaling are otherwise almost never taken by random data."""
    lib, L = _lib()
    g = torch.Generator().manual_seed(5)
    S, H, T, D = 7, 8, 1152, 64
    lens = torch.tensor([1152, 1151, 1, 64, 65, 129, 700], dtype=torch.int32)
    q = (torch.randn(S, H, T, D, generator=g) * 0.25).half()
    k = (torch.randn(S, H, T, D, generator=g) * 2).half()
    v = torch.randn(S, H, T, D, generator=g).half()
    # sequences 0 and 6: keys aligned with the queries and growing with the key index: scores climb by ~60 nats over the sequence
    ramp = torch.linspace(0.0, 6.0, T)[None, :, None]
    for s_ in (0, 6):
        base = torch.randn(H, 1, D, generator=g)
        base = base / base.norm(dim=-1, keepdim=True)
        q[s_] = (base * 2.5 + torch.randn(H, T, D, generator=g) * 0.05).half()
        k[s_] = (base * 4.0 * ramp + torch.randn(H, T, D, generator=g) * 0.5).half()
    ref = _attn_ref(q, k, v, lens, chunk)
    qd, kd = q.cuda(), k.cuda()
    vtd = v.transpose(2, 3).contiguous().cuda()
    out = torch.zeros(S, T, H * D, dtype=torch.float16, device="cuda")
    lens_d = lens.cuda()
    lib.check(L.cv2_op_flash_attn(_s(), lib.ptr(qd), lib.ptr(kd), lib.ptr(vtd), lib.ptr(out), lib.ptr(lens_d), 0, S, H, T, chunk))
    torch.cuda.synchronize()
    o = out.float().cpu()
    assert bool(torch.isfinite(o).all())
    for s_ in range(S):
        n = int(lens[s_])
        err = float((o[s_, :n] - ref[s_, :n]).abs().max())
        assert err < 8e-3, (s_, n, err)
        assert float(o[s_, n:].abs().max()) == 0.0 if n < T else True


@pytest.mark.parametrize("chunk,T,Tal,len1", [(0, 100, 128, 61), (25, 100, 128, 61), (0, 300, 384, 170), (50, 300, 384, 257)])
def test_rel_attention(chunk, T, Tal, len1):
    import token2wav_oracle as O
    lib, L = _lib()
    g = torch.Generator().manual_seed(5)
    S = 2
    lens = torch.tensor([T, len1], dtype=torch.int32)
    x = torch.randn(S, T, 512, generator=g)
    p = {n: torch.randn(512, 512, generator=g) / math.sqrt(512) for n in
         ("linear_q.weight", "linear_k.weight", "linear_v.weight", "linear_out.weight", "linear_pos.weight")}
    for n in ("linear_q.bias", "linear_k.bias", "linear_v.bias", "linear_out.bias"):
        p[n] = torch.randn(512, generator=g) * 0.1
    p["pos_bias_u"] = torch.randn(8, 64, generator=g) * 0.1
    p["pos_bias_v"] = torch.randn(8, 64, generator=g) * 0.1
    p["linear_out.weight"] = torch.eye(512)
    p["linear_out.bias"] = torch.zeros(512)
    F = torch.nn.functional
    q = F.linear(x, p["linear_q.weight"], p["linear_q.bias"]).view(S, T, 8, 64)
    k = F.linear(x, p["linear_k.weight"], p["linear_k.bias"]).view(S, T, 8, 64)
    v = F.linear(x, p["linear_v.weight"], p["linear_v.bias"]).view(S, T, 8, 64)

    def heads(a):                                   # [S,T,8,64] -> [S,8,Tal,64] 16-bit
        o = torch.zeros(S, 8, Tal, 64)
        o[:, :, :T] = a.permute(0, 2, 1, 3)
        return o.half()
    qu, qv, kk = heads((q + p["pos_bias_u"]) * 0.125), heads((q + p["pos_bias_v"]) * 0.125), heads(k)
    vt = heads(v).transpose(2, 3).contiguous()     # [S,8,64,Tal]
    # table by relative position for Tmax = Tal: row (rel + Tal - 1)
    R_alloc = (2 * Tal - 1 + 127) // 128 * 128
    pos = F.linear(O.rel_pos_table(Tal)[0], p["linear_pos.weight"])          # rows: rel = Tal-1 ... -(Tal-1)
    pos16 = torch.zeros(R_alloc, 512, dtype=torch.float16)
    pos16[:2 * Tal - 1] = torch.flip(pos, [0]).half()                        # -> row r <-> rel = r - (Tal-1)
    out = torch.zeros(S, Tal, 512, dtype=torch.float16, device="cuda")
    keep = [qu.cuda(), qv.cuda(), kk.cuda(), vt.cuda(), pos16.cuda(), lens.cuda()]
    lib.check(L.cv2_op_rel_attn(_s(), lib.ptr(keep[0]), lib.ptr(keep[1]), lib.ptr(keep[2]), lib.ptr(keep[3]), lib.ptr(keep[4]),
                                lib.ptr(out), lib.ptr(keep[5]), 0, S, Tal, Tal, R_alloc, chunk))
    torch.cuda.synchronize()
    for s in range(S):
        n = int(lens[s])
        valid = torch.ones(1, 1, n, dtype=torch.bool)
        mask = O.attention_mask(valid, chunk)
        ref = O.rel_mha(x[s:s + 1, :n], mask, O.rel_pos_table(n), p)
        err = float((out[s, :n].float().cpu() - ref[0]).abs().max())
        assert err < 5e-3, (s, err)


def test_source_stft_and_istft():
    import token2wav_oracle as O
    lib, L = _lib()
    g = torch.Generator().manual_seed(7)
    B, T = 2, 6
    lens = torch.tensor([6, 4], dtype=torch.int32)
    src = torch.tanh(torch.randn(B, 480 * T, generator=g))
    F_alloc = 120 * T + 1 + 63
    out = torch.zeros(B, F_alloc, 18, device="cuda")
    src_d, lens_d = src.cuda(), lens.cuda()
    lib.check(L.cv2_op_source_stft(_s(), lib.ptr(src_d), T, lib.ptr(lens_d), lib.ptr(out), F_alloc, B))
    cp = torch.randn(B, F_alloc, 18, generator=g) * 1.5
    cp[:, :, 0] += 4.0      # push one bin through the 1e2 clip
    wav = torch.zeros(B, 480 * T, device="cuda")
    cp_d = cp.cuda()
    lib.check(L.cv2_op_istft(_s(), lib.ptr(cp_d), F_alloc, lib.ptr(lens_d), T, lib.ptr(wav), B))
    torch.cuda.synchronize()
    for b in range(B):
        n = int(lens[b])
        ref = O.source_stft(src[b:b + 1, :480 * n])[0].t()                     # [F, 18]
        Fr = ref.shape[0]
        assert float((out[b, :Fr].cpu() - ref).abs().max()) < 1e-5
        assert float(out[b, Fr:].abs().max()) == 0.0
        refw = O.istft_head(cp[b:b + 1, :Fr].transpose(1, 2))[0]
        assert refw.shape[0] == 480 * n
        assert float((wav[b, :480 * n].cpu() - refw).abs().max()) < 2e-4
        assert float(refw.abs().max()) == pytest.approx(0.99)


def test_nsf_source():
    import token2wav_oracle as O
    lib, L = _lib()
    g = torch.Generator().manual_seed(11)
    B, T = 2, 40
    f0 = torch.rand(B, T, generator=g) * 400
    f0[:, 5:9] = 3.0          # unvoiced frames
    f0[1, 30:] = 0.0
    noise = torch.randn(B, 480 * T, 9, generator=g)
    lw = torch.randn(1, 9, generator=g) * 0.3
    lb = torch.randn(1, generator=g) * 0.1
    sd = {"m_source.l_linear.weight": lw, "m_source.l_linear.bias": lb}
    ref = O.nsf_source(sd, f0, noise)[:, 0]
    ph = torch.zeros(B, T, 9, device="cuda")
    src = torch.zeros(B, 480 * T, device="cuda")
    keep = [f0.cuda(), noise.cuda(), lw.reshape(-1).cuda(), lb.cuda()]   # device pointers must outlive the launch
    lib.check(L.cv2_op_nsf_source(_s(), lib.ptr(keep[0]), T, None, lib.ptr(keep[1]), 0, lib.ptr(keep[2]), lib.ptr(keep[3]),
                                  lib.ptr(ph), lib.ptr(src), B))
    torch.cuda.synchronize()
    err = (src.cpu() - ref).abs()
    refd = ref.double().numpy()
    snr = 10 * np.log10((refd ** 2).sum() / ((refd - src.cpu().double().numpy()) ** 2).sum())
    print("nsf source max err", float(err.max()), "snr", snr)
    assert snr > 50
