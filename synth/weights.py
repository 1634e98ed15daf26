"""SYNTHETIC DATA (tests + bench) -- deterministic random-init weights of the CosyVoice2-0.5B-EU
token2wav architecture (flow 1121 tensors / 112.5 M params, hift 328 tensors /
20.8 M params).

There is no checkpoint offline (weights live on HF hi-paris/CosyVoice2-0.5B-EU),
and north_star asks for random-init weights of the architecture.  The schema
(name -> shape) in schema_flow.json / schema_hift.json was dumped from
`module.state_dict()` of the reference modules built by oracle/ref_shims.py
(constructor args = cosy_repo/examples/libritts/cosyvoice2/conf/cosyvoice2.yaml:39-112).

Values are drawn per tensor from a numpy Philox stream keyed by sha256(seed:name),
so they do not depend on torch version, tensor order or platform, and the GPU box
regenerates exactly the tensors the golden vectors were produced with.

Fixture rescaling (SURVEY.md 8c "Random-init fixture caveats"): with plain random
init the F0 predictor emits ~0.05 Hz (0 % voiced), exp(conv_post) never reaches
the 1e2 clip and the audio never reaches the +-0.99 clamp.  So:
  * f0_predictor.classifier.weight x F0_GAIN, bias = F0_BIAS  -> f0 spans ~0..400 Hz
    with both voiced (f0 > 10) and unvoiced frames,
  * conv_post effective weight x POST_GAIN                  -> magnitudes pass 1e2,
  * Snake alpha ~ lognormal(0, 0.5)                          -> alpha != 1.
"""
import hashlib
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

F0_GAIN = 300.0     # classifier pre-abs output ~ N(-0.16, 0.37) on flow mels -> f0 = |300 p + 60| ~ |N(12, 112)| Hz
F0_BIAS = 60.0
HIFT_GAIN = 0.8     # scale of every HiFT conv weight relative to unit-gain init (keeps stage activations O(1))
POST_GAIN = 3.6     # extra gain on conv_post (log-magnitude / phase head)
POST_MAG_BIAS = -1.0  # mean of the 9 log-magnitude biases


def load_schema(which):
    with open(os.path.join(_HERE, f"schema_{which}.json")) as f:
        return json.load(f)


def _rng(seed, name):
    key = int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:8], "little")
    return np.random.Generator(np.random.Philox(key=key))


def _uniform(rng, shape, bound):
    return ((rng.random(shape, dtype=np.float32) * 2.0 - 1.0) * np.float32(bound)).astype(np.float32)


def _normal(rng, shape, std, mean=0.0):
    return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std) + np.float32(mean)).astype(np.float32)


def _is_layernorm(name):
    # LayerNorm parameters in the flow schema (1-D `.weight` whose sibling is not a conv/linear)
    keys = (".norm1.", ".norm3.", ".norm_ff.", ".norm_mha.", "after_norm.", ".block.2.", "embed.out.1.")
    return any(k in name for k in keys)


def make_flow_state(seed=1234):
    """-> dict name -> np.float32 array, loadable by the reference flow.load_state_dict."""
    out = {}
    for name, shape in load_schema("flow").items():
        rng = _rng(seed, "flow." + name)
        shape = tuple(shape)
        if name == "input_embedding.weight":
            v = _normal(rng, shape, 1.0)
        elif "pos_bias_" in name:
            v = _normal(rng, shape, 0.1)
        elif _is_layernorm(name):
            v = _normal(rng, shape, 0.1, mean=1.0 if name.endswith("weight") else 0.0)
        elif name.endswith(".bias"):
            v = _normal(rng, shape, 0.05)
        else:  # Linear / Conv weight: PyTorch-default-like U(-1/sqrt(fan_in), 1/sqrt(fan_in))
            fan_in = int(np.prod(shape[1:]))
            v = _uniform(rng, shape, 1.0 / np.sqrt(fan_in))
        out[name] = v
    return out


def make_hift_state(seed=1234):
    """-> dict name -> np.float32 array, loadable by the reference hift.load_state_dict
    (torch>=2.1 weight-norm parametrization names: original0 = g, original1 = v)."""
    schema = load_schema("hift")
    out = {}
    for name, shape in schema.items():
        rng = _rng(seed, "hift." + name)
        shape = tuple(shape)
        if name.endswith(".alpha"):
            v = np.exp(_normal(rng, shape, 0.5)).astype(np.float32)
        elif name.endswith("original1") or (name.endswith(".weight") and len(shape) >= 2):
            if name.startswith("ups."):
                fan_in = shape[0] * shape[2] / {"ups.0": 8, "ups.1": 5, "ups.2": 3}[name[:5]]
            else:
                fan_in = int(np.prod(shape[1:]))
            gain = 1.0 if name.startswith("f0_predictor.") else HIFT_GAIN
            v = _uniform(rng, shape, gain * np.sqrt(3.0) / np.sqrt(fan_in))  # std = gain/sqrt(fan_in)
        elif name.endswith("original0"):
            v = None  # filled below from ||v||
        elif name.endswith(".bias"):
            v = _normal(rng, shape, 0.05)
        else:
            raise KeyError(name)
        out[name] = v
    for name in schema:
        if name.endswith("original0"):
            vname = name[:-1] + "1"
            vv = out[vname]
            norm = np.sqrt((vv.astype(np.float64) ** 2).sum(axis=tuple(range(1, vv.ndim)), keepdims=True))
            jitter = 1.0 + 0.1 * _rng(seed, "hift." + name).standard_normal(norm.shape)
            out[name] = (norm * jitter).astype(np.float32)
    out["f0_predictor.classifier.weight"] = (out["f0_predictor.classifier.weight"] * np.float32(F0_GAIN)).astype(np.float32)
    out["f0_predictor.classifier.bias"] = np.full((1,), F0_BIAS, np.float32)
    out["conv_post.bias"][:9] += np.float32(POST_MAG_BIAS)
    out["conv_post.parametrizations.weight.original0"] = (
        out["conv_post.parametrizations.weight.original0"] * np.float32(POST_GAIN)).astype(np.float32)
    return out


def to_torch(state):
    import torch
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in state.items()}


# ----------------------------------------------------------------------------------------
# synthetic inputs (SURVEY.md 8d "Synthetic inputs")
# ----------------------------------------------------------------------------------------
def make_utterance(n_tokens, n_prompt=75, seed=0):
    """One synthetic utterance: speech tokens, prompt tokens, prompt mel, x-vector, NSF noise."""
    rng = _rng(seed, f"utt.{n_tokens}.{n_prompt}")
    token = rng.integers(0, 6561, size=(1, n_tokens), dtype=np.int64).astype(np.int32)
    prompt_token = rng.integers(0, 6561, size=(1, n_prompt), dtype=np.int64).astype(np.int32)
    prompt_feat = np.clip(_normal(rng, (1, 2 * n_prompt, 80), 2.0, mean=-5.0), -11.5, 2.0).astype(np.float32)
    embedding = _normal(rng, (1, 192), 1.0)
    return dict(token=token, prompt_token=prompt_token, prompt_feat=prompt_feat, embedding=embedding)


def make_nsf_noise(n_samples, seed=0):
    """Gaussian noise [1, n_samples, 9] injected into SineGen2 (generator.py:334) in both
    the oracle and the engine so hift parity is deterministic."""
    return _normal(_rng(seed, f"nsf.{n_samples}"), (1, n_samples, 9), 1.0)


def cfm_rand_noise():
    """CausalConditionalCFM.rand_noise (flow_matching.py:195-198): torch.randn([1,80,15000])
    right after set_all_random_seed(0), CPU generator."""
    import torch
    g = torch.Generator(device="cpu")
    g.manual_seed(0)
    return torch.randn([1, 80, 50 * 300], generator=g)
