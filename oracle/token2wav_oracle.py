"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU (torch fp32, functional, batch-1) restatement of the reference's token2wav
algorithm: speech tokens + prompt mel + x-vector -> 24 kHz waveform.  It exists only
to check the CUDA engine; only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it.  The product path
(cosyvoice2_eu_b200/) never imports anything from oracle/.

Pinning: tests/test_oracle_vs_reference.py runs this file against the unmodified
reference imported from /root/reference (container only) and
tests/test_oracle_golden.py checks it against tests/golden/*.npz, which
oracle/make_golden.py generated from the REFERENCE's own modules.  The one boundary
whose source is not under /root/reference is diffusers==0.29.0 Attention/GELU
(requirements.txt:5): "parity unpinned" there -- restated from its published
semantics (see oracle/ref_shims.py).

All `path:line` citations are relative to /root/reference/cosy_repo/; CV = cosyvoice,
MT = third_party/Matcha-TTS/matcha.

Weights come in as the reference's own state_dict names (Appendix A of SURVEY.md).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------
# integer / boolean bookkeeping (must be bit-exact)                      CV/utils/mask.py
# ------------------------------------------------------------------------------------------


def pad_mask(lengths, max_len=0):
    """True where t >= len[b].  CV/utils/mask.py:239-265."""
    lengths = torch.as_tensor(lengths).reshape(-1).long()
    n = int(max_len) if max_len > 0 else int(lengths.max())
    return torch.arange(n)[None, :] >= lengths[:, None]


def chunk_mask(size, chunk):
    """ret[i, j] = j < (i // chunk + 1) * chunk; num_left_chunks is ignored.  CV/utils/mask.py:127-158."""
    i = torch.arange(size)
    return i[None, :] < ((i // chunk + 1) * chunk)[:, None]


def attention_mask(valid, chunk):
    """valid: bool [B,1,T] (True = real frame).  chunk > 0 -> AND with the block-causal chunk
    mask, rows that end up all-False are forced all-True.  CV/utils/mask.py:161-236."""
    if chunk > 0:
        m = valid & chunk_mask(valid.shape[-1], chunk)[None]
    else:
        m = valid
    dead = m.sum(-1) == 0
    if bool(dead.any()):
        m = m.clone()
        m[dead] = True
    return m


def stream_schedule(n_tokens, n_prompt, hop=25, lookahead=3, all_tokens_ready=False):
    """Chunk schedule of CosyVoice2Model.tts(stream=True), CV/cli/model.py:351-381.
    Returns list of (n_tokens_visible, token_offset, finalize).

    The loop computes `this_token_hop_len` at the top of an iteration (model.py:356) and its break test (model.py:369) reuses
    that value after `token_offset` has moved.  That only matters in the iteration that emits the FIRST chunk (hop + pad), and
    only if the LLM has already finished by then: `all_tokens_ready=True` (vc_job, model.py:141-143, or an LLM faster than the
    first chunk) reproduces it -- the loop then finalizes as soon as fewer than hop + pad + lookahead tokens remain.  With
    tokens still arriving (`False`) every later iteration recomputes the hop and the two tests agree.
    tests/golden/tts_schedule.npz holds the reference's own call sequences for the first case."""
    pad = int(math.ceil(n_prompt / hop) * hop - n_prompt)
    calls, off = [], 0
    while True:
        this_hop = hop + pad if off == 0 else hop
        if n_tokens - off >= this_hop + lookahead:
            calls.append((off + this_hop + lookahead, off, False))
            off += this_hop
            if all_tokens_ready and n_tokens - off < this_hop + lookahead:      # model.py:369 with the stale hop
                break
        else:
            break
    calls.append((n_tokens, off, True))
    return calls


# ------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def fold_weight_norm(sd, prefix):
    """w = g * v / ||v|| over all dims but 0 (torch parametrizations.weight_norm, dim=0);
    CV/hifigan/generator.py:26-29 selects the parametrization form.  Falls back to a plain
    `.weight` when the layer is not weight-normed."""
    g = sd.get(prefix + ".parametrizations.weight.original0")
    if g is None:
        return sd[prefix + ".weight"]
    v = sd[prefix + ".parametrizations.weight.original1"]
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / norm)


# ------------------------------------------------------------------------------------------
# UpsampleConformerEncoder                         CV/transformer/upsample_encoder.py:243-306
# ------------------------------------------------------------------------------------------


def rel_pos_table(T, d=512):
    """EspnetRelPositionalEncoding slice for a length-T input: rows are relative positions
    T-1, ..., 0, ..., -(T-1).  CV/transformer/embedding.py:228-253, 292-296."""
    pos = torch.arange(T - 1, -T, -1, dtype=torch.float32)[:, None]
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * -(math.log(10000.0) / d))
    pe = torch.zeros(2 * T - 1, d)
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe[None]


def rel_shift(x):
    """[B,H,T,2T-1] -> [B,H,T,T], out[i,j] = x[i, j - i + T - 1].  CV/transformer/attention.py:225-247."""
    T = x.shape[2]
    idx = torch.arange(T)[None, :] - torch.arange(T)[:, None] + T - 1
    return x.gather(-1, idx.expand(x.shape[0], x.shape[1], T, T))


def rel_mha(x, mask, pos_emb, p, heads=8):
    """RelPositionMultiHeadedAttention.forward, CV/transformer/attention.py:249-330 with
    forward_attention :71-120 (masked_fill(-inf) -> softmax -> masked_fill(0))."""
    B, T, D = x.shape
    dk = D // heads
    q = F.linear(x, p["linear_q.weight"], p["linear_q.bias"]).view(B, T, heads, dk)
    k = F.linear(x, p["linear_k.weight"], p["linear_k.bias"]).view(B, T, heads, dk).transpose(1, 2)
    v = F.linear(x, p["linear_v.weight"], p["linear_v.bias"]).view(B, T, heads, dk).transpose(1, 2)
    pe = F.linear(pos_emb, p["linear_pos.weight"]).view(1, -1, heads, dk).transpose(1, 2)
    qu = (q + p["pos_bias_u"]).transpose(1, 2)
    qv = (q + p["pos_bias_v"]).transpose(1, 2)
    ac = qu @ k.transpose(-2, -1)
    bd = rel_shift(qv @ pe.transpose(-2, -1))
    scores = (ac + bd) / math.sqrt(dk)
    dead = ~mask[:, None]                      # [B,1,T|1,T]
    scores = scores.masked_fill(dead, float("-inf"))
    attn = torch.softmax(scores, dim=-1).masked_fill(dead, 0.0)
    o = (attn @ v).transpose(1, 2).reshape(B, T, D)
    return F.linear(o, p["linear_out.weight"], p["linear_out.bias"])


def conformer_layer(x, mask, pos_emb, p):
    """ConformerEncoderLayer (no macaron, no conv module), CV/transformer/encoder_layer.py:160-236;
    LN eps 1e-12 (:145-146); FFN = w_2(SiLU(w_1(x))), positionwise_feed_forward.py:55."""
    h = _ln(x, p["norm_mha.weight"], p["norm_mha.bias"], 1e-12)
    x = x + rel_mha(h, mask, pos_emb, _sub(p, "self_attn."))
    h = _ln(x, p["norm_ff.weight"], p["norm_ff.bias"], 1e-12)
    h = F.linear(F.silu(F.linear(h, p["feed_forward.w_1.weight"], p["feed_forward.w_1.bias"])),
                 p["feed_forward.w_2.weight"], p["feed_forward.w_2.bias"])
    return x + h


def linear_embed(x, p):
    """LinearNoSubsampling + EspnetRelPositionalEncoding: Linear -> LN(1e-5) -> x*sqrt(d).
    CV/transformer/subsampling.py:83-113, embedding.py:256-270."""
    x = _ln(F.linear(x, p["out.0.weight"], p["out.0.bias"]), p["out.1.weight"], p["out.1.bias"], 1e-5)
    return x * math.sqrt(x.shape[-1])


def encoder_forward(sd, xs, xs_len, context=None, streaming=False):
    """UpsampleConformerEncoder.forward.  xs [1,T,512] (token embeddings), context [1,3,512] or None.
    sd = flow state dict restricted to `encoder.`.  Returns h [1,2T,512]."""
    B, T, _ = xs.shape
    valid = ~pad_mask(xs_len, T)[:, None, :]
    x = linear_embed(xs, _sub(sd, "embed."))
    pos = rel_pos_table(T)
    cm = attention_mask(valid, 25 if streaming else 0)                       # :285
    # PreLookaheadLayer :81-102 -- right context = real lookahead tokens (embedded) or zeros
    y = x.transpose(1, 2)
    if context is not None and context.shape[1] > 0:
        ctx = linear_embed(context, _sub(sd, "embed.")).transpose(1, 2)
        y = torch.cat([y, ctx], dim=2)
        y = F.pad(y, (0, 3 - ctx.shape[2]))
    else:
        y = F.pad(y, (0, 3))
    y = F.leaky_relu(F.conv1d(y, sd["pre_lookahead_layer.conv1.weight"], sd["pre_lookahead_layer.conv1.bias"]))
    y = F.conv1d(F.pad(y, (2, 0)), sd["pre_lookahead_layer.conv2.weight"], sd["pre_lookahead_layer.conv2.bias"])
    x = y.transpose(1, 2) + x
    for i in range(6):
        x = conformer_layer(x, cm, pos, _sub(sd, f"encoders.{i}."))
    # Upsample1D :59-63 -- nearest x2, left pad 4, conv k5
    y = x.transpose(1, 2).repeat_interleave(2, dim=2)
    y = F.conv1d(F.pad(y, (4, 0)), sd["up_layer.conv.weight"], sd["up_layer.conv.bias"])
    x = y.transpose(1, 2)
    T2 = x.shape[1]
    valid2 = ~pad_mask(torch.as_tensor(xs_len) * 2, T2)[:, None, :]
    x = linear_embed(x, _sub(sd, "up_embed."))
    pos2 = rel_pos_table(T2)
    cm2 = attention_mask(valid2, 50 if streaming else 0)                     # :298
    for i in range(4):
        x = conformer_layer(x, cm2, pos2, _sub(sd, f"up_encoders.{i}."))
    return _ln(x, sd["after_norm.weight"], sd["after_norm.bias"], 1e-5)


# ------------------------------------------------------------------------------------------
# CausalConditionalDecoder (CFM estimator)                        CV/flow/decoder.py:405-494
# ------------------------------------------------------------------------------------------


def time_embedding(sd, t):
    """SinusoidalPosEmb(320, scale 1000) -> Linear -> SiLU -> Linear.
    MT/models/components/decoder.py:14-29, 73-117."""
    half = 160
    freq = torch.exp(torch.arange(half).float() * -(math.log(10000) / (half - 1)))
    e = 1000.0 * t[:, None] * freq[None]
    e = torch.cat([e.sin(), e.cos()], dim=-1)
    e = F.silu(F.linear(e, sd["time_mlp.linear_1.weight"], sd["time_mlp.linear_1.bias"]))
    return F.linear(e, sd["time_mlp.linear_2.weight"], sd["time_mlp.linear_2.bias"])


def causal_block(x, m, p):
    """CausalBlock1D: (x*m) -> left-pad 2 -> conv k3 -> LayerNorm over C -> Mish -> *m.
    CV/flow/decoder.py:65-78."""
    y = F.conv1d(F.pad(x * m, (2, 0)), p["block.0.weight"], p["block.0.bias"])
    y = _ln(y.transpose(1, 2), p["block.2.weight"], p["block.2.bias"], 1e-5).transpose(1, 2)
    return F.mish(y) * m


def causal_resnet(x, m, temb, p):
    """ResnetBlock1D.forward, MT/models/components/decoder.py:56-61, with causal blocks."""
    h = causal_block(x, m, _sub(p, "block1."))
    h = h + F.linear(F.mish(temb), p["mlp.1.weight"], p["mlp.1.bias"])[:, :, None]
    h = causal_block(h, m, _sub(p, "block2."))
    return h + F.conv1d(x * m, p["res_conv.weight"], p["res_conv.bias"])


def transformer_block(x, bias, p, heads=8):
    """BasicTransformerBlock (MT/models/components/transformer.py:243-316) with diffusers 0.29.0
    Attention (bias-free q/k/v, scale 1/sqrt(64), additive mask, out-proj + bias) and
    FeedForward = Linear -> exact GELU -> Linear (:83-134)."""
    B, T, C = x.shape
    h = _ln(x, p["norm1.weight"], p["norm1.bias"], 1e-5)
    q = F.linear(h, p["attn1.to_q.weight"]).view(B, T, heads, -1).transpose(1, 2)
    k = F.linear(h, p["attn1.to_k.weight"]).view(B, T, heads, -1).transpose(1, 2)
    v = F.linear(h, p["attn1.to_v.weight"]).view(B, T, heads, -1).transpose(1, 2)
    s = (q @ k.transpose(-2, -1)) / math.sqrt(q.shape[-1]) + bias[:, None]
    a = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, T, -1)
    x = x + F.linear(a, p["attn1.to_out.0.weight"], p["attn1.to_out.0.bias"])
    h = _ln(x, p["norm3.weight"], p["norm3.bias"], 1e-5)
    h = F.gelu(F.linear(h, p["ff.net.0.proj.weight"], p["ff.net.0.proj.bias"]))
    return x + F.linear(h, p["ff.net.2.weight"], p["ff.net.2.bias"])


def estimator_forward(sd, x, mask, mu, t, spks, cond, streaming=False):
    """CausalConditionalDecoder.forward.  x/mu/cond [B,80,T], mask [B,1,T] float, t [B], spks [B,80].
    sd = flow state dict restricted to `decoder.estimator.`."""
    temb = time_embedding(sd, t)
    T = x.shape[-1]
    h = torch.cat([x, mu, spks[:, :, None].expand(-1, -1, T), cond], dim=1)       # :425-431
    valid = mask.bool()
    am = attention_mask(valid, 50 if streaming else 0)
    if not streaming:
        am = am.expand(-1, T, -1)                                                 # .repeat(1, T, 1) :441
    bias = (1.0 - am.float()) * -1.0e10                                           # mask_to_bias common.py:160-168

    def group(h, prefix, n_tfm=4):
        h = causal_resnet(h, mask, temb, _sub(sd, prefix + "0."))
        y = h.transpose(1, 2)
        for j in range(n_tfm):
            y = transformer_block(y, bias, _sub(sd, f"{prefix}1.{j}."))
        return y.transpose(1, 2)

    h = group(h, "down_blocks.0.")
    skip = h
    h = F.conv1d(F.pad(h * mask, (2, 0)), sd["down_blocks.0.2.weight"], sd["down_blocks.0.2.bias"])
    for i in range(12):
        h = group(h, f"mid_blocks.{i}.")
    h = torch.cat([h, skip], dim=1)
    h = group(h, "up_blocks.0.")
    h = F.conv1d(F.pad(h * mask, (2, 0)), sd["up_blocks.0.2.weight"], sd["up_blocks.0.2.bias"])
    h = causal_block(h, mask, _sub(sd, "final_block."))
    return F.conv1d(h * mask, sd["final_proj.weight"], sd["final_proj.bias"]) * mask


def t_schedule(n_steps=10):
    """Cosine schedule, CV/flow/flow_matching.py:222-224."""
    return 1 - torch.cos(torch.linspace(0, 1, n_steps + 1) * 0.5 * math.pi)


def solve_euler(sd_est, z, mu, mask, spks, cond, n_steps=10, cfg=0.7, streaming=False, estimator=None):
    """ConditionalCFM.solve_euler, CV/flow/flow_matching.py:71-123: batch-2 CFG (row 0 cond, row 1 zeros),
    dphi = (1+cfg) v_c - cfg v_u, x += dt*dphi, dt from the RUNNING t."""
    ts = t_schedule(n_steps)
    t, dt = ts[0:1], ts[1] - ts[0]
    x = z
    est = estimator or (lambda *a, **k: estimator_forward(sd_est, *a, **k))
    for step in range(1, n_steps + 1):
        xin = torch.cat([x, x], 0)
        min_ = torch.cat([mask, mask], 0)
        muin = torch.cat([mu, torch.zeros_like(mu)], 0)
        tin = torch.cat([t, t], 0)
        spin = torch.cat([spks, torch.zeros_like(spks)], 0)
        cin = torch.cat([cond, torch.zeros_like(cond)], 0)
        v = est(xin, min_, muin, tin, spin, cin, streaming=streaming)
        d = (1.0 + cfg) * v[0:1] - cfg * v[1:2]
        x = x + dt * d
        t = t + dt
        if step < n_steps:
            dt = ts[step + 1] - t
    return x.float()


# ------------------------------------------------------------------------------------------
# CausalMaskedDiffWithXvec.inference                                  CV/flow/flow.py:235-283
# ------------------------------------------------------------------------------------------


def flow_inference(sd, rand_noise, token, prompt_token, prompt_feat, embedding, streaming=False, finalize=True,
                   return_intermediates=False):
    """token int [1,N], prompt_token int [1,P], prompt_feat f32 [1,2P,80], embedding f32 [1,192]
    -> mel f32 [1,80,T_gen].  rand_noise = CausalConditionalCFM.rand_noise [1,80,15000]."""
    emb = F.linear(F.normalize(embedding, dim=1), sd["spk_embed_affine_layer.weight"], sd["spk_embed_affine_layer.bias"])
    tok = torch.cat([prompt_token, token], dim=1)
    tok_len = torch.tensor([tok.shape[1]])
    x = F.embedding(torch.clamp(tok, min=0).long(), sd["input_embedding.weight"])
    x = x * (~pad_mask(tok_len, tok.shape[1]))[:, :, None].float()
    if finalize:
        h = encoder_forward(_sub(sd, "encoder."), x, tok_len, None, streaming)
    else:
        h = encoder_forward(_sub(sd, "encoder."), x[:, :-3], tok_len, x[:, -3:], streaming)      # :262-263
    mel_len1 = prompt_feat.shape[1]
    mel_len2 = h.shape[1] - mel_len1
    mu = F.linear(h, sd["encoder_proj.weight"], sd["encoder_proj.bias"]).transpose(1, 2).contiguous()
    T = mel_len1 + mel_len2
    cond = torch.zeros(1, 80, T)
    cond[:, :, :mel_len1] = prompt_feat.transpose(1, 2)
    mask = torch.ones(1, 1, T)
    z = rand_noise[:, :, :T].clone()
    feat = solve_euler(_sub(sd, "decoder.estimator."), z, mu, mask, emb, cond, streaming=streaming)
    mel = feat[:, :, mel_len1:]
    assert mel.shape[2] == mel_len2
    if return_intermediates:
        return mel, dict(mu=mu, spks=emb, cond=cond, h=h)
    return mel


# ------------------------------------------------------------------------------------------
# HiFTGenerator                                                 CV/hifigan/generator.py:520-582
# ------------------------------------------------------------------------------------------

UP_RATES, UP_KERNELS = (8, 5, 3), (16, 11, 7)
RES_KERNELS, RES_DILATIONS = (3, 7, 11), (1, 3, 5)
SRC_RES_KERNELS = (7, 7, 11)
HOP_TOTAL = 480          # prod(upsample_rates) * istft hop (generator.py:431,437)
N_FFT, HOP = 16, 4


def hann16():
    """scipy get_window('hann', 16, fftbins=True) = periodic Hann.  generator.py:487."""
    n = np.arange(16, dtype=np.float64)
    return torch.from_numpy((0.5 * (1.0 - np.cos(2.0 * np.pi * n / 16.0))).astype(np.float32))


def f0_predict(sd, mel):
    """ConvRNNF0Predictor.forward, CV/hifigan/f0_predictor.py:55-58.  mel [B,80,T] -> f0 [B,T]."""
    x = mel
    for i in (0, 2, 4, 6, 8):
        x = F.elu(F.conv1d(x, fold_weight_norm(sd, f"condnet.{i}"), sd[f"condnet.{i}.bias"], padding=1))
    return torch.abs(F.linear(x.transpose(1, 2), sd["classifier.weight"], sd["classifier.bias"]).squeeze(-1))


def nsf_source(sd, f0, noise):
    """SourceModuleHnNSF2 / SineGen2 (generator.py:375-389, 261-283, 314-339), closed form of
    SURVEY.md Appendix D.  f0 [B,T] (Hz, frame rate), noise [B,480T,9] ~ N(0,1) (injected; the
    reference draws torch.randn_like at :334).  rand_ini (:270-272) is a no-op at scale 480 and is
    not modelled.  Returns source [B,1,480T].
    The frame-rate cumsum follows ATen's CPU cumsum on fp32 (fp64 accumulate, fp32 outputs)."""
    B, T = f0.shape
    harm = torch.arange(1, 10, dtype=torch.float32)
    f0s = f0[:, None, :].repeat_interleave(HOP_TOTAL, dim=2).transpose(1, 2)           # f0_upsamp :437,574
    fn = f0s * harm[None, None, :]
    rad = (fn / 24000.0) % 1
    rad = F.interpolate(rad.transpose(1, 2), scale_factor=1 / HOP_TOTAL, mode="linear").transpose(1, 2)
    phase = torch.cumsum(rad, dim=1) * 2 * np.pi
    phase = F.interpolate(phase.transpose(1, 2) * HOP_TOTAL, scale_factor=HOP_TOTAL, mode="linear").transpose(1, 2)
    sine = torch.sin(phase) * 0.1
    uv = (f0s > 10).float()
    amp = uv * 0.003 + (1 - uv) * 0.1 / 3
    sine = sine * uv + amp * noise
    src = torch.tanh(F.linear(sine, sd["m_source.l_linear.weight"], sd["m_source.l_linear.bias"]))
    return src.transpose(1, 2)


def source_stft(s):
    """torch.stft(n_fft 16, hop 4, periodic hann, center/reflect) -> cat(real, imag) [B,18,L/4+1].
    generator.py:504-510, 521-522."""
    spec = torch.stft(s, N_FFT, HOP, N_FFT, window=hann16(), return_complex=True)
    return torch.cat([spec.real, spec.imag], dim=1)


def snake(x, alpha):
    """x + sin^2(alpha x)/(alpha + 1e-9), alpha per channel.  CV/transformer/activation.py:73-84."""
    a = alpha[None, :, None]
    return x + (1.0 / (a + 1e-9)) * torch.sin(x * a) ** 2


def resblock(sd, prefix, x, k):
    """ResBlock.forward, generator.py:94-101."""
    for j, d in enumerate(RES_DILATIONS):
        xt = snake(x, sd[f"{prefix}.activations1.{j}.alpha"])
        xt = F.conv1d(xt, fold_weight_norm(sd, f"{prefix}.convs1.{j}"), sd[f"{prefix}.convs1.{j}.bias"],
                      dilation=d, padding=(k * d - d) // 2)
        xt = snake(xt, sd[f"{prefix}.activations2.{j}.alpha"])
        xt = F.conv1d(xt, fold_weight_norm(sd, f"{prefix}.convs2.{j}"), sd[f"{prefix}.convs2.{j}.bias"],
                      padding=(k - 1) // 2)
        x = xt + x
    return x


def istft_head(x):
    """conv_post output [B,18,F] -> waveform [B,4(F-1)]: exp / sin, clip 1e2, istft, clamp +-0.99.
    generator.py:546-551, 512-518."""
    mag = torch.clip(torch.exp(x[:, :9]), max=1e2)
    ph = torch.sin(x[:, 9:])
    spec = torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))
    y = torch.istft(spec, N_FFT, HOP, N_FFT, window=hann16())
    return torch.clamp(y, -0.99, 0.99)


def hift_decode(sd, mel, s, return_intermediates=False):
    """HiFTGenerator.decode, generator.py:520-552.  mel [B,80,T], s [B,1,480T]."""
    inter = {}
    s_stft = source_stft(s.squeeze(1))
    inter["s_stft"] = s_stft
    x = F.conv1d(mel, fold_weight_norm(sd, "conv_pre"), sd["conv_pre.bias"], padding=3)
    down_rates = (15, 3, 1)
    for i in range(3):
        x = F.leaky_relu(x, 0.1)
        u, k = UP_RATES[i], UP_KERNELS[i]
        x = F.conv_transpose1d(x, fold_weight_norm(sd, f"ups.{i}"), sd[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if i == 2:
            x = F.pad(x, (1, 0), mode="reflect")
        r = down_rates[i]
        if r == 1:
            si = F.conv1d(s_stft, sd[f"source_downs.{i}.weight"], sd[f"source_downs.{i}.bias"])
        else:
            si = F.conv1d(s_stft, sd[f"source_downs.{i}.weight"], sd[f"source_downs.{i}.bias"], stride=r, padding=r // 2)
        si = resblock(sd, f"source_resblocks.{i}", si, SRC_RES_KERNELS[i])
        x = x + si
        inter[f"fused{i}"] = x
        acc = None
        for j in range(3):
            y = resblock(sd, f"resblocks.{i * 3 + j}", x, RES_KERNELS[j])
            acc = y if acc is None else acc + y
        x = acc / 3
        inter[f"stage{i}"] = x
    x = F.leaky_relu(x)                                                       # default slope 0.01, :545
    x = F.conv1d(x, fold_weight_norm(sd, "conv_post"), sd["conv_post.bias"], padding=3)
    inter["conv_post"] = x
    y = istft_head(x)
    if return_intermediates:
        return y, inter
    return y


def hift_inference(sd, mel, cache_source=None, noise=None, return_intermediates=False):
    """HiFTGenerator.inference, generator.py:570-582 -> (speech [B,480T], source [B,1,480T])."""
    f0 = f0_predict(_sub(sd, "f0_predictor."), mel)
    if noise is None:
        noise = torch.randn(mel.shape[0], mel.shape[2] * HOP_TOTAL, 9)
    s = nsf_source(sd, f0, noise)
    if cache_source is not None and cache_source.shape[2] != 0:
        s = s.clone()
        s[:, :, :cache_source.shape[2]] = cache_source
    if return_intermediates:
        y, inter = hift_decode(sd, mel, s, True)
        inter["f0"] = f0
        return y, s, inter
    return hift_decode(sd, mel, s), s


# ------------------------------------------------------------------------------------------
# CosyVoice2Model.token2wav                                          CV/cli/model.py:300-334
# ------------------------------------------------------------------------------------------

MEL_CACHE_LEN = 8
SOURCE_CACHE_LEN = 8 * 480


def fade_in_out(new, old, window):
    """CV/utils/common.py:142-150 -- float64 numpy hamming window applied to f32 tensors."""
    n = window.shape[0] // 2
    out = new.clone()
    w = torch.from_numpy(window)
    out[..., :n] = (new[..., :n] * w[:n] + old[..., -n:] * w[n:]).to(out.dtype)
    return out


class OracleToken2Wav:
    """Stateful restatement of CosyVoice2Model.token2wav incl. the per-uuid hift cache."""

    def __init__(self, flow_sd, hift_sd, rand_noise):
        self.flow_sd, self.hift_sd, self.rand_noise = flow_sd, hift_sd, rand_noise
        self.hift_cache_dict = {}
        self.speech_window = np.hamming(2 * SOURCE_CACHE_LEN)

    @torch.inference_mode()
    def token2wav(self, token, prompt_token, prompt_feat, embedding, token_offset, uuid, stream=False, finalize=False,
                  speed=1.0, noise=None):
        mel = flow_inference(self.flow_sd, self.rand_noise, token, prompt_token, prompt_feat, embedding,
                             streaming=stream, finalize=finalize)
        mel = mel[:, :, token_offset * 2:]
        cache = self.hift_cache_dict.get(uuid)
        if cache is not None:
            mel = torch.cat([cache["mel"], mel], dim=2)
            src_cache = cache["source"]
        else:
            src_cache = torch.zeros(1, 1, 0)
        if not finalize:
            speech, source = hift_inference(self.hift_sd, mel, src_cache, noise)
            if cache is not None:
                speech = fade_in_out(speech, cache["speech"], self.speech_window)
            self.hift_cache_dict[uuid] = dict(mel=mel[:, :, -MEL_CACHE_LEN:], source=source[:, :, -SOURCE_CACHE_LEN:],
                                              speech=speech[:, -SOURCE_CACHE_LEN:])
            speech = speech[:, :-SOURCE_CACHE_LEN]
        else:
            if speed != 1.0:
                assert cache is None
                mel = F.interpolate(mel, size=int(mel.shape[2] / speed), mode="linear")
            speech, source = hift_inference(self.hift_sd, mel, src_cache, noise)
            if cache is not None:
                speech = fade_in_out(speech, cache["speech"], self.speech_window)
        return speech
