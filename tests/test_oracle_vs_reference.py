"""Container-only pin (skipped wherever /root/reference is absent, i.e. on the GPU box): the committed golden vectors are what
the UNMODIFIED reference computes, and the oracle agrees with the live reference on an input no golden file holds.
Uses oracle/ref_shims.py (import shims for omegaconf / conformer / matcha.utils / diffusers)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/cosy_repo"), reason="the reference tree only exists in the build container")


@pytest.fixture(scope="module")
def reference():
    import ref_shims
    import weights
    torch.manual_seed(0)
    flow, hift = ref_shims.build_reference_modules()
    flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
    hift.load_state_dict(weights.to_torch(weights.make_hift_state()))
    return flow, hift


def _call(flow, u, streaming=False, finalize=True):
    t = torch.from_numpy
    n, p = u["token"].shape[1], u["prompt_token"].shape[1]
    with torch.inference_mode():
        mel, _ = flow.inference(token=t(u["token"]), token_len=torch.tensor([n], dtype=torch.int32), prompt_token=t(u["prompt_token"]),
                                prompt_token_len=torch.tensor([p], dtype=torch.int32), prompt_feat=t(u["prompt_feat"]),
                                prompt_feat_len=torch.tensor([2 * p], dtype=torch.int32), embedding=t(u["embedding"]),
                                streaming=streaming, finalize=finalize)
    return mel


def test_golden_tiny_is_what_the_reference_computes(reference, golden):
    import weights
    g = golden("tiny")
    u = weights.make_utterance(int(g["n_tok"]), int(g["n_prompt"]), int(g["seed"]))
    assert np.array_equal(_call(reference[0], u).numpy(), g["mel"])
    assert np.array_equal(_call(reference[0], u, streaming=True, finalize=False).numpy(), g["mel_stream_nonfinal"])


def test_oracle_equals_live_reference_on_a_fresh_input(reference, fixture_weights):
    import token2wav_oracle as O
    import weights
    u = weights.make_utterance(23, 7, 4321)
    t = torch.from_numpy
    for streaming, finalize in ((False, True), (True, False)):
        ref = _call(reference[0], u, streaming, finalize)
        with torch.inference_mode():
            got = O.flow_inference(fixture_weights[0], weights.cfm_rand_noise(), t(u["token"]), t(u["prompt_token"]), t(u["prompt_feat"]),
                                   t(u["embedding"]), streaming=streaming, finalize=finalize)
        assert tuple(got.shape) == tuple(ref.shape)
        assert float((got - ref).abs().max()) < 1e-4
