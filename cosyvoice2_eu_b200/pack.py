"""Weight packing: reference state_dict -> the named tensors the C-ABI engine consumes.

Input is exactly what the reference loads (`flow.load_state_dict`, `hift.load_state_dict` with the
`generator.` prefix stripped, cosyvoice/cli/model.py:85-90; schema in SURVEY.md Appendix A).
Transformations (all host-side, once per model):
  * weight-norm folded: w = g * v / ||v||  (cosyvoice/hifigan/generator.py:26-29; the reference's own
    remove_weight_norm is broken, generator.py:498);
  * conv weights [Cout, Cin, k] -> tap-major [Cout, k * Cin_pad] (implicit-GEMM K axis), Cin padded to 64;
  * q/k/v projections concatenated to one [3*inner, dim] matrix;
  * ConvTranspose1d [Cin, Cout, k] (stride s) -> phase-concatenated [s*Cout, ceil(k/s) * Cin];
  * sqrt(512) of EspnetRelPositionalEncoding folded into the embed LayerNorm's gamma/beta;
  * MMA operands stored as fp16 (bf16 operands miss the 1e-2 mel budget, fp16 has the same
    tcgen05 rate), everything an epilogue reads (bias, LayerNorm, Snake alpha) and the F0 predictor stay fp32.
"""
import math

import torch

OP_DTYPE = torch.float16
F0_WSCALE = 256.0   # F0-predictor split weights are stored x256 (engine: acc_scale = 1/256)
CFG_RATE = 0.7   # inference_cfg_rate (cosyvoice2.yaml:75); the engine falls back to the unfused update for any other rate


def _conv_w(w, cin_pad=None):
    """[Cout, Cin, k] -> [Cout, k*Cin_pad] tap-major, 16-bit."""
    cout, cin, k = w.shape
    cp = cin_pad or ((cin + 63) // 64 * 64)
    out = torch.zeros(cout, k, cp, dtype=torch.float32)
    out[:, :, :cin] = w.permute(0, 2, 1)
    return out.reshape(cout, k * cp).to(OP_DTYPE).contiguous()


def _lin_w(w):
    return w.to(OP_DTYPE).contiguous()


def _f32(x):
    return x.detach().to(torch.float32).contiguous()


def fold_weight_norm(sd, prefix):
    g = sd.get(prefix + ".parametrizations.weight.original0")
    if g is None:
        g = sd.get(prefix + ".weight_g")
        v = sd.get(prefix + ".weight_v")
        if g is None:
            return sd[prefix + ".weight"].float()
    else:
        v = sd[prefix + ".parametrizations.weight.original1"]
    v = v.float()
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g.float() / norm)


def pack_flow(sd):
    """sd: flow.state_dict() (CPU tensors) -> dict name -> CPU tensor."""
    o = {}
    o["flow.embedding"] = _f32(sd["input_embedding.weight"])
    o["flow.spk.w"] = _f32(sd["spk_embed_affine_layer.weight"])
    o["flow.spk.b"] = _f32(sd["spk_embed_affine_layer.bias"])
    xs = math.sqrt(512.0)
    for src, dst in (("encoder.embed", "enc.embed"), ("encoder.up_embed", "enc.up_embed")):
        o[dst + ".w"] = _lin_w(sd[src + ".out.0.weight"])
        o[dst + ".b"] = _f32(sd[src + ".out.0.bias"])
        o[dst + ".ln_g"] = _f32(sd[src + ".out.1.weight"] * xs)
        o[dst + ".ln_b"] = _f32(sd[src + ".out.1.bias"] * xs)
    for c in ("conv1", "conv2"):
        o[f"enc.pre.{c}.w"] = _conv_w(sd[f"encoder.pre_lookahead_layer.{c}.weight"])
        o[f"enc.pre.{c}.b"] = _f32(sd[f"encoder.pre_lookahead_layer.{c}.bias"])
    o["enc.up.conv.w"] = _conv_w(sd["encoder.up_layer.conv.weight"])
    o["enc.up.conv.b"] = _f32(sd["encoder.up_layer.conv.bias"])
    for src, dst, n in (("encoder.encoders", "enc.layers", 6), ("encoder.up_encoders", "enc.up_layers", 4)):
        for i in range(n):
            s, d = f"{src}.{i}", f"{dst}.{i}"
            a = s + ".self_attn"
            # column blocks (q + u) | (q + v) | k | v: the Transformer-XL biases pos_bias_u / pos_bias_v (attention.py:308-311)
            # are added to q before the two score matmuls, so they fold into the bias of two copies of linear_q
            wq, bq = sd[a + ".linear_q.weight"], sd[a + ".linear_q.bias"].float()
            o[d + ".qkv.w"] = _lin_w(torch.cat([wq, wq, sd[a + ".linear_k.weight"], sd[a + ".linear_v.weight"]], 0))
            o[d + ".qkv.b"] = _f32(torch.cat([bq + sd[a + ".pos_bias_u"].float().reshape(-1), bq + sd[a + ".pos_bias_v"].float().reshape(-1),
                                              sd[a + ".linear_k.bias"].float(), sd[a + ".linear_v.bias"].float()], 0))
            o[d + ".pos.w"] = _lin_w(sd[a + ".linear_pos.weight"])
            o[d + ".o.w"] = _lin_w(sd[a + ".linear_out.weight"])
            o[d + ".o.b"] = _f32(sd[a + ".linear_out.bias"])
            o[d + ".ff1.w"] = _lin_w(sd[s + ".feed_forward.w_1.weight"])
            o[d + ".ff1.b"] = _f32(sd[s + ".feed_forward.w_1.bias"])
            o[d + ".ff2.w"] = _lin_w(sd[s + ".feed_forward.w_2.weight"])
            o[d + ".ff2.b"] = _f32(sd[s + ".feed_forward.w_2.bias"])
            for ln in ("norm_mha", "norm_ff"):
                o[f"{d}.ln_{ln[5:]}_g"] = _f32(sd[f"{s}.{ln}.weight"])
                o[f"{d}.ln_{ln[5:]}_b"] = _f32(sd[f"{s}.{ln}.bias"])
    o["enc.after_norm_g"] = _f32(sd["encoder.after_norm.weight"])
    o["enc.after_norm_b"] = _f32(sd["encoder.after_norm.bias"])
    o["enc.proj.w"] = _lin_w(sd["encoder_proj.weight"])
    o["enc.proj.b"] = _f32(sd["encoder_proj.bias"])

    e = "decoder.estimator."
    o["est.time.w1"] = _f32(sd[e + "time_mlp.linear_1.weight"])
    o["est.time.b1"] = _f32(sd[e + "time_mlp.linear_1.bias"])
    o["est.time.w2"] = _f32(sd[e + "time_mlp.linear_2.weight"])
    o["est.time.b2"] = _f32(sd[e + "time_mlp.linear_2.bias"])
    groups = ["down_blocks.0"] + [f"mid_blocks.{i}" for i in range(12)] + ["up_blocks.0"]
    for r, g in enumerate(groups):
        s, d = f"{e}{g}.0", f"est.res.{r}"
        o[d + ".mlp_w"] = _f32(sd[s + ".mlp.1.weight"])
        o[d + ".mlp_b"] = _f32(sd[s + ".mlp.1.bias"])
        for blk, c, ln in (("block1", "c1", "ln1"), ("block2", "c2", "ln2")):
            o[f"{d}.{c}.w"] = _conv_w(sd[f"{s}.{blk}.block.0.weight"])
            o[f"{d}.{c}.b"] = _f32(sd[f"{s}.{blk}.block.0.bias"])
            o[f"{d}.{ln}_g"] = _f32(sd[f"{s}.{blk}.block.2.weight"])
            o[f"{d}.{ln}_b"] = _f32(sd[f"{s}.{blk}.block.2.bias"])
        o[d + ".res.w"] = _lin_w(sd[s + ".res_conv.weight"][:, :, 0])
        o[d + ".res.b"] = _f32(sd[s + ".res_conv.bias"])
        for j in range(4):
            s, d = f"{e}{g}.1.{j}", f"est.tfm.{r}.{j}"
            o[d + ".ln1_g"] = _f32(sd[s + ".norm1.weight"])
            o[d + ".ln1_b"] = _f32(sd[s + ".norm1.bias"])
            # the 1/sqrt(64) of the attention scores is folded into the q rows (a power of two: exact), so the QKV epilogue -- which is
            # issue bound -- does not multiply
            o[d + ".qkv.w"] = _lin_w(torch.cat([sd[s + ".attn1.to_q.weight"] * 0.125, sd[s + ".attn1.to_k.weight"], sd[s + ".attn1.to_v.weight"]], 0))
            o[d + ".o.w"] = _lin_w(sd[s + ".attn1.to_out.0.weight"])
            o[d + ".o.b"] = _f32(sd[s + ".attn1.to_out.0.bias"])
            o[d + ".ln3_g"] = _f32(sd[s + ".norm3.weight"])
            o[d + ".ln3_b"] = _f32(sd[s + ".norm3.bias"])
            o[d + ".ff1.w"] = _lin_w(sd[s + ".ff.net.0.proj.weight"])
            o[d + ".ff1.b"] = _f32(sd[s + ".ff.net.0.proj.bias"])
            o[d + ".ff2.w"] = _lin_w(sd[s + ".ff.net.2.weight"])
            o[d + ".ff2.b"] = _f32(sd[s + ".ff.net.2.bias"])
    for src, dst in (("down_blocks.0.2", "est.down_conv"), ("up_blocks.0.2", "est.up_conv"), ("final_block.block.0", "est.final.c")):
        o[dst + ".w"] = _conv_w(sd[e + src + ".weight"])
        o[dst + ".b"] = _f32(sd[e + src + ".bias"])
    o["est.final.ln_g"] = _f32(sd[e + "final_block.block.2.weight"])
    o["est.final.ln_b"] = _f32(sd[e + "final_block.block.2.bias"])
    o["est.proj.w"] = _lin_w(sd[e + "final_proj.weight"][:, :, 0])
    o["est.proj.b"] = _f32(sd[e + "final_proj.bias"])
    # CFG combine folded into the projection (flow_matching.py:116: dphi = (1+r) v_cond - r v_uncond, r = inference_cfg_rate):
    # K-concatenated [(1+r) W | -r W] over the (cond, uncond) pair of rows; the bias is unchanged ((1+r) b - r b = b)
    w = sd[e + "final_proj.weight"][:, :, 0].float()
    o["est.proj_cfg.w"] = _lin_w(torch.cat([(1.0 + CFG_RATE) * w, -CFG_RATE * w], 1))
    o["est.proj_cfg.b"] = _f32(sd[e + "final_proj.bias"])
    return o


UP_RATES, UP_KERNELS = (8, 5, 3), (16, 11, 7)


def _convT_w(wt, stride):
    """ConvTranspose1d weight [Cin, Cout, k] -> phase-concatenated GEMM weight [stride*Cout, q*Cin]:
    row r*Cout + c, column q*Cin + ci holds wt[ci, c, r + stride*q] (0 where r + stride*q >= k)."""
    cin, cout, k = wt.shape
    q = (k + stride - 1) // stride
    out = torch.zeros(stride, cout, q, cin, dtype=torch.float32)
    for r in range(stride):
        for qq in range(q):
            j = r + stride * qq
            if j < k:
                out[r, :, qq, :] = wt[:, :, j].t()
    return out.reshape(stride * cout, q * cin).to(OP_DTYPE).contiguous()


def pack_hift(sd):
    """sd: hift.state_dict() (CPU tensors) -> dict name -> CPU tensor."""
    o = {}
    for l, idx in enumerate((0, 2, 4, 6, 8)):
        w = fold_weight_norm(sd, f"f0_predictor.condnet.{idx}")
        o[f"f0.c{l}.w"] = _f32(w.permute(2, 1, 0))                     # [3][Cin][512]
        o[f"f0.c{l}.b"] = _f32(sd[f"f0_predictor.condnet.{idx}.bias"])
    # split-precision tensor-core form of the same convs: W = W_hi + W_lo (16-bit halves); per output channel the K axis is
    # 6 "taps" x [A_hi | A_lo] blocks: taps 0-2 hold [W_hi | W_hi] (conv taps -1, 0, +1), taps 3-5 hold [W_lo | 0]
    for l, idx in enumerate((0, 2, 4, 6, 8)):
        w = fold_weight_norm(sd, f"f0_predictor.condnet.{idx}")          # [512, Cin, 3]
        cout, cin, k = w.shape
        cp = (cin + 127) // 128 * 128 if cin < 512 else cin
        w = w * F0_WSCALE                                                 # keeps W_lo (~2^-11 |W|) a normal fp16 number
        hi = w.to(OP_DTYPE).float()
        lo = (w - hi).to(OP_DTYPE).float()
        blk = torch.zeros(cout, 6, 2, cp, dtype=torch.float32)
        for j in range(3):
            blk[:, j, 0, :cin] = hi[:, :, j]
            blk[:, j, 1, :cin] = hi[:, :, j]
            blk[:, 3 + j, 0, :cin] = lo[:, :, j]
        o[f"f0.t{l}.w"] = blk.reshape(cout, 6 * 2 * cp).to(OP_DTYPE).contiguous()
        o[f"f0.t{l}.b"] = _f32(sd[f"f0_predictor.condnet.{idx}.bias"])
    o["f0.cls.w"] = _f32(sd["f0_predictor.classifier.weight"].reshape(-1))
    o["f0.cls.b"] = _f32(sd["f0_predictor.classifier.bias"].reshape(-1))
    o["hift.src.lw"] = _f32(sd["m_source.l_linear.weight"].reshape(-1))
    o["hift.src.lb"] = _f32(sd["m_source.l_linear.bias"].reshape(-1))
    o["hift.conv_pre.w"] = _conv_w(fold_weight_norm(sd, "conv_pre"), cin_pad=128)
    o["hift.conv_pre.b"] = _f32(sd["conv_pre.bias"])
    for i in range(3):
        o[f"hift.ups.{i}.w"] = _convT_w(fold_weight_norm(sd, f"ups.{i}"), UP_RATES[i])
        o[f"hift.ups.{i}.b"] = _f32(sd[f"ups.{i}.bias"].repeat(UP_RATES[i]))
        o[f"hift.sd.{i}.w"] = _f32(sd[f"source_downs.{i}.weight"].permute(2, 1, 0))   # [k][18][C]
        o[f"hift.sd.{i}.b"] = _f32(sd[f"source_downs.{i}.bias"])
    blocks = [(f"source_resblocks.{i}", f"hift.srb.{i}") for i in range(3)] + [(f"resblocks.{n}", f"hift.rb.{n}") for n in range(9)]
    for s, d in blocks:
        for j in range(3):
            o[f"{d}.c1.{j}.w"] = _conv_w(fold_weight_norm(sd, f"{s}.convs1.{j}"))
            o[f"{d}.c1.{j}.b"] = _f32(sd[f"{s}.convs1.{j}.bias"])
            o[f"{d}.c2.{j}.w"] = _conv_w(fold_weight_norm(sd, f"{s}.convs2.{j}"))
            o[f"{d}.c2.{j}.b"] = _f32(sd[f"{s}.convs2.{j}.bias"])
            o[f"{d}.a1.{j}"] = _f32(sd[f"{s}.activations1.{j}.alpha"])
            o[f"{d}.a2.{j}"] = _f32(sd[f"{s}.activations2.{j}.alpha"])
    o["hift.conv_post.w"] = _conv_w(fold_weight_norm(sd, "conv_post"))
    o["hift.conv_post.b"] = _f32(sd["conv_post.bias"])
    return o
