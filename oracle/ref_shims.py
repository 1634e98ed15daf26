"""TEST INFRASTRUCTURE ONLY -- import shims that let the *unmodified* reference
(/root/reference, read-only) be imported in the build container so that the
oracle restatement (oracle/token2wav_oracle.py) can be pinned against it and
golden vectors can be generated (oracle/make_golden.py).

/root/reference does NOT exist on the GPU box: nothing in `-m gpu` tests,
smoke() or bench.py may import this module.  Only oracle/make_golden.py and
the container-only pinning test (tests/test_oracle_vs_reference.py, skipped
when /root/reference is absent) use it.

Four third-party modules the reference imports are absent here (SURVEY.md
section 8c): omegaconf, conformer, matcha.utils(.pylogger) and diffusers.  The
first three are import-time only.  `diffusers==0.29.0` (requirements.txt:5)
carries real arithmetic (Attention + GELU used by
third_party/Matcha-TTS/matcha/models/components/transformer.py:5-14,110,196);
the shim restates its published semantics (AttnProcessor2_0: bias-free q/k/v
Linear, F.scaled_dot_product_attention with additive mask, out Linear + bias;
GELU = Linear + exact erf GELU).  The reference holds no test that pins this
boundary => "parity unpinned" at the diffusers boundary.
"""
import logging
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = "/root/reference/cosy_repo"
MATCHA_ROOT = "/root/reference/cosy_repo/third_party/Matcha-TTS"


class DictConfig(dict):
    """omegaconf.DictConfig stand-in: dict with attribute access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class _Attention(nn.Module):
    """diffusers 0.29.0 `Attention` restricted to what the estimator uses
    (self-attention, AttnProcessor2_0, no norm/rescale/residual inside)."""

    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, **kw):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(query_dim, inner, bias=bias)
        self.to_v = nn.Linear(query_dim, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(dropout)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        b, t, _ = hidden_states.shape
        h = self.heads
        q = self.to_q(hidden_states).view(b, t, h, -1).transpose(1, 2)
        k = self.to_k(hidden_states).view(b, t, h, -1).transpose(1, 2)
        v = self.to_v(hidden_states).view(b, t, h, -1).transpose(1, 2)
        if attention_mask is not None:
            # prepare_attention_mask: [B, Tq, Tk] -> repeat_interleave(heads) -> [B, H, Tq, Tk]
            attention_mask = attention_mask.unsqueeze(1).expand(b, h, attention_mask.shape[-2], attention_mask.shape[-1])
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask, dropout_p=0.0, is_causal=False)
        o = o.transpose(1, 2).reshape(b, t, -1)
        o = self.to_out[0](o)
        o = self.to_out[1](o)
        return o


class _GELU(nn.Module):
    def __init__(self, dim_in, dim_out, approximate="none", bias=True):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out, bias=bias)
        self.approximate = approximate

    def forward(self, x):
        return F.gelu(self.proj(x), approximate=self.approximate)


class _Unused(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()


def _get_activation(name):
    name = name.lower()
    return {"silu": nn.SiLU(), "swish": nn.SiLU(), "mish": nn.Mish(), "gelu": nn.GELU(), "relu": nn.ReLU()}[name]


def install():
    """Install the shims into sys.modules and put the reference on sys.path."""
    if "cosyvoice" in sys.modules and getattr(sys.modules.get("omegaconf"), "_b200_shim", False):
        return
    for p in (REF_ROOT, MATCHA_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m._b200_shim = True
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    mod("omegaconf", DictConfig=DictConfig)
    mod("conformer", ConformerBlock=_Unused)
    mod("diffusers")
    mod("diffusers.models")
    mod("diffusers.models.activations", get_activation=_get_activation)
    mod("diffusers.models.attention", GEGLU=_Unused, GELU=_GELU, AdaLayerNorm=_Unused, AdaLayerNormZero=_Unused,
        ApproximateGELU=_Unused)
    mod("diffusers.models.attention_processor", Attention=_Attention)
    mod("diffusers.models.lora", LoRACompatibleLinear=nn.Linear)
    mod("diffusers.utils")
    mod("diffusers.utils.torch_utils", maybe_allow_in_graph=lambda c: c)
    # matcha.utils pulls hydra / lightning at import time; only get_pylogger is used.
    import importlib.util
    import os
    matcha = types.ModuleType("matcha")
    matcha.__path__ = [os.path.join(MATCHA_ROOT, "matcha")]
    sys.modules["matcha"] = matcha
    mu = mod("matcha.utils")
    mu.__path__ = []
    mod("matcha.utils.pylogger", get_pylogger=lambda name=__name__: logging.getLogger(name))
    mu.get_pylogger = sys.modules["matcha.utils.pylogger"].get_pylogger


def build_reference_modules():
    """Construct the reference flow + hift modules with the hyper-parameters of
    cosy_repo/examples/libritts/cosyvoice2/conf/cosyvoice2.yaml:39-112
    (SURVEY.md Appendix C). Returns (flow, hift) in eval mode, random init."""
    install()
    from cosyvoice.flow.decoder import CausalConditionalDecoder
    from cosyvoice.flow.flow import CausalMaskedDiffWithXvec
    from cosyvoice.flow.flow_matching import CausalConditionalCFM
    from cosyvoice.hifigan.f0_predictor import ConvRNNF0Predictor
    from cosyvoice.hifigan.generator import HiFTGenerator
    from cosyvoice.transformer.upsample_encoder import UpsampleConformerEncoder

    enc = UpsampleConformerEncoder(input_size=512, output_size=512, attention_heads=8, linear_units=2048, num_blocks=6,
                                   dropout_rate=0.1, positional_dropout_rate=0.1, attention_dropout_rate=0.1,
                                   normalize_before=True, input_layer='linear', pos_enc_layer_type='rel_pos_espnet',
                                   selfattention_layer_type='rel_selfattn', use_cnn_module=False, macaron_style=False,
                                   static_chunk_size=25)
    est = CausalConditionalDecoder(in_channels=320, out_channels=80, channels=[256], dropout=0.0, attention_head_dim=64,
                                   n_blocks=4, num_mid_blocks=12, num_heads=8, act_fn='gelu', static_chunk_size=50,
                                   num_decoding_left_chunks=-1)
    cfm = CausalConditionalCFM(in_channels=240, n_spks=1, spk_emb_dim=80,
                               cfm_params=DictConfig(sigma_min=1e-6, solver='euler', t_scheduler='cosine',
                                                     training_cfg_rate=0.2, inference_cfg_rate=0.7, reg_loss_type='l1'),
                               estimator=est)
    flow = CausalMaskedDiffWithXvec(input_size=512, output_size=80, spk_embed_dim=192, output_type='mel', vocab_size=6561,
                                    input_frame_rate=25, only_mask_loss=True, token_mel_ratio=2, pre_lookahead_len=3,
                                    encoder=enc, decoder=cfm).eval()
    hift = HiFTGenerator(in_channels=80, base_channels=512, nb_harmonics=8, sampling_rate=24000, nsf_alpha=0.1,
                         nsf_sigma=0.003, nsf_voiced_threshold=10, upsample_rates=[8, 5, 3],
                         upsample_kernel_sizes=[16, 11, 7], istft_params={'n_fft': 16, 'hop_len': 4},
                         resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5]] * 3,
                         source_resblock_kernel_sizes=[7, 7, 11], source_resblock_dilation_sizes=[[1, 3, 5]] * 3,
                         lrelu_slope=0.1, audio_limit=0.99,
                         f0_predictor=ConvRNNF0Predictor(num_class=1, in_channels=80, cond_channels=512)).eval()
    return flow, hift


def build_reference_model(flow, hift):
    """CosyVoice2Model with a stub LLM (never touched on the token2wav path)."""
    install()
    from cosyvoice.cli.model import CosyVoice2Model
    return CosyVoice2Model(llm=nn.Identity(), flow=flow, hift=hift, fp16=False)
