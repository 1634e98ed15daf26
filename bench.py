#!/usr/bin/env python
"""bench.py -- token2wav throughput on B200 (BASELINE.json metric: audio-seconds per second; RTF at batch 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path (flow.inference + hift.inference, i.e. token2wav offline) over one batch of
synthetic utterances per GPU.  Workload at every N: BASELINE.json configs[2] per GPU -- 64 variable-length
utterances, durations U[4, 20] s (N = round(25 d) tokens, 75-token / 150-frame prompt), weak scaling (every rank gets
its own 64-utterance shard of the corpus, as configs[4] shards by utterance; the only collective is the final gather).
configs[1] (single 10 s utterance, batch 1) is timed as well and reported under "rtf_batch1".

`value`   : generated audio-seconds / second with all inputs resident in HBM (device timed, max over ranks).
`e2e`     : the same through the public API (B200Token2Wav.token2wav_batch) with host inputs in pinned memory,
            H2D of tokens / prompt mel / x-vectors and D2H of the waveforms inside the timed region (+ NCCL gather of
            lengths and audio to rank 0 when N > 1).
`roofline`: dominant kernel family by device time (CUDA events around every launch of the family, recorded on the
            launching stream during the timed region): algorithmic FLOPs / family time vs the measured dense bf16 peak.
`cpu_baseline` / --impl reference: the oracle port of the reference's CPU token2wav (torch fp32, all host threads) on a
            bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PROMPT = 75
BATCH = 64
FAMILIES = ["gemm_tap<64>", "gemm_tap<128>", "gemm_tap<256>", "flash_attn", "rel_attn", "f0_conv_f32", "nsf_source", "source_stft",
            "source_down", "istft", "ffn_fused"]
# gemm_tap<256> launches are additionally split by epilogue specialisation (engine families 11 + spec)
G256 = ["generic", "qkv_split", "res+ln_emit", "gelu", "conv+ln+mish+temb", "conv+ln+mish+res+ln_emit", "res+plain_emit", "silu",
        "res", "out32", "plain_emit", "snake", "res+snake_emit", "res_sum"]
NFAM = len(FAMILIES) + len(G256)


def workload(rank, batch=BATCH):
    """configs[2]: durations U[4,20] s -> tokens.  Weak scaling fixes the work PER GPU, so every rank draws the same 64
    lengths (the token / prompt / x-vector CONTENT is seeded per rank); round 1 drew different lengths per rank, which folded
    the rank-to-rank variance of a 64-utterance sample (+-4 % in cost) into the scaling efficiency."""
    rng = np.random.Generator(np.random.Philox(key=1000))
    dur = rng.uniform(4.0, 20.0, size=batch)
    return [int(round(25 * d)) for d in dur]


def algorithmic_flops(n_tokens, n_prompt=N_PROMPT):
    """Unpadded algorithmic FLOPs (multiply-add = 2) per kernel family for a list of utterances; layer dims from
    cosyvoice2.yaml:39-112, totals agree with SURVEY.md 8(d) (2.43 TFLOP for the 8 s utterance of configs[0])."""
    f = dict.fromkeys(FAMILIES, 0.0)
    for n in n_tokens:
        tt = n + n_prompt          # token-rate frames
        T = 2 * tt                 # mel frames seen by the estimator
        tg = 2 * n                 # generated mel frames (hift)
        est_lin = 132161536.0
        ffn = 56 * 2 * 2 * 256 * 1024.0          # FF1 + FF2 of the 56 transformer blocks, per frame per CFG row
        fused = os.environ.get("CV2_NO_FFN_FUSION") is None
        # the attention out-projection (56 x [512 -> 256]) is chained into the FFN kernel on the 2-SM path (big launches)
        outp = 56 * 2 * 512 * 256.0 if (fused and os.environ.get("CV2_NO_OUTPROJ_CHAIN") is None) else 0.0
        f["gemm_tap<256>"] += 20 * T * (est_lin - 40960.0 - (ffn if fused else 0.0) - outp)
        f["ffn_fused"] += 20 * T * ((ffn if fused else 0.0) + outp)
        f["gemm_tap<128>"] += 20 * T * 40960.0
        f["flash_attn"] += 20 * T * 114688.0 * T
        f["gemm_tap<256>"] += tt * (4.194304e6 + 6 * 7.340032e6) + T * (3.227648e6 - 81920.0 + 4 * 7.340032e6)
        f["gemm_tap<128>"] += T * 81920.0
        f["rel_attn"] += tt * 6 * 4096.0 * tt + T * 4 * 8192.0 * tt
        f["gemm_tap<256>"] += tg * (573440.0 + 4194304.0 + 8 * 336 * 65536.0)
        f["gemm_tap<128>"] += tg * (8 * 720896.0 + 40 * 336 * 16384.0)
        f["gemm_tap<64>"] += tg * (40 * 114688.0 + 120 * 384 * 4096.0 + 120 * 16128.0)
        f["source_down"] += tg * (8 * 276480.0 + 40 * 27648.0 + 120 * 2304.0)
        f["gemm_tap<256>" if os.environ.get("CV2_F0_SPLIT") else "f0_conv_f32"] += tg * 6.54e6   # F0 predictor
    return f


def sample_clocks(stop, out):
    q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    idx = os.environ.get("LOCAL_RANK", "0")
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", idx, f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                             stdout=subprocess.PIPE, text=True)
    except Exception:
        return
    while not stop.is_set():
        line = p.stdout.readline()
        if not line:
            break
        out.append(line.strip())
    p.terminate()


def clocks_summary(lines):
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        c = [x.strip() for x in ln.split(",")]
        try:
            sm.append(float(c[0]))
            mx = max(mx, float(c[1]))
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        except Exception:
            continue
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
            "samples": len(sm)}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p, "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


_CPU_ENGINE = {}


def cpu_reference_run(n_tok, threads=None):
    """Oracle port of the reference CPU token2wav on one utterance; returns (seconds, audio_seconds)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import token2wav_oracle as O
    from synth import weights
    if threads:
        torch.set_num_threads(threads)
    if "eng" not in _CPU_ENGINE:
        fs, hs = weights.to_torch(weights.make_flow_state()), weights.to_torch(weights.make_hift_state())
        _CPU_ENGINE["eng"] = O.OracleToken2Wav(fs, hs, weights.cfm_rand_noise())
    eng = _CPU_ENGINE["eng"]
    u = {k: torch.from_numpy(v) for k, v in weights.make_utterance(n_tok, N_PROMPT, seed=7).items()}
    eng.hift_cache_dict["b"] = None
    t0 = time.perf_counter()
    wav = eng.token2wav(u["token"], u["prompt_token"], u["prompt_feat"], u["embedding"], 0, "b", finalize=True)
    dt = time.perf_counter() - t0
    return dt, wav.shape[1] / 24000.0


def workload_config(world, batch=BATCH):
    """`config` of the JSON line: the same dictionary in both arms (the reference arm runs a bounded SAMPLE of it)."""
    n0 = workload(0, batch)
    return {"workload": "BASELINE configs[2] per GPU: 64 utterances, durations U[4,20] s (100-500 tokens) + 75-token "
                        "prompt, offline token2wav, 10 Euler steps with CFG, length-sorted ragged batch",
            "utterances_per_gpu": len(n0), "audio_seconds_per_step_per_gpu": sum(2 * n * 480 for n in n0) / 24000.0,
            "parallelism": f"utterance-sharded dp{world}",
            "l2": "per-step working set (GBs of activations) is far larger than the 126 MB L2; no explicit flush",
            "weights": "random-init CosyVoice2-0.5B-EU architecture (synth/weights.py)"}


def main_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the Python reference cannot
    travel to the GPU box), all host threads, bounded sample = one median-length (12 s) utterance of the workload per step.
    Under torchrun only rank 0 works; the other ranks exit at once."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)           # torchrun exports OMP_NUM_THREADS=1: undo that for the CPU arm
    n_tok = 300
    times, audio = [], 0.0
    for i in range(args.warmup + args.steps):
        dt, a = cpu_reference_run(n_tok, cores)
        if i >= args.warmup:
            times.append(dt)
            audio += a
    total = sum(times)
    val = audio / total
    sample = "one 12 s utterance (300 tokens + 75-token prompt) of the workload per step, oracle port of the reference CPU " \
             "token2wav, torch fp32, all host threads"
    line = {"impl": "reference", "metric": "token2wav_audio_seconds_per_second", "value": val, "unit": "audio-s/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total / max(len(times), 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus, args.batch),
            "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def kernel_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` captures of this very command, as
    extracted into profiles/kernel_traffic.json by profiles/extract_traffic.py (tracked; names the capture it came from)."""
    try:
        with open(os.path.join(ROOT, "profiles", "kernel_traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def main():
    if os.environ.get("CV2_DBG_SKIP_EPI"):
        # the GEMM epilogue measurement switch skips work: a bench line taken with it would be invalid
        sys.exit("bench.py refuses to run with CV2_DBG_SKIP_EPI set (work would be skipped inside the timed region); use "
                 "profiles/prof_flow.py for that experiment")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--corpus", type=int, default=4096, help="utterances of the configs[4] corpus leg (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-legs", action="store_true", help="skip batch-1 / streaming / prompt-mel legs (quick runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)

    import datetime
    import ctypes as C
    import torch
    import torch.distributed as dist
    from cosyvoice2_eu_b200 import B200Flow, B200HiFT, B200Token2Wav, shard
    from cosyvoice2_eu_b200.scheduler import chunk_schedule
    from synth import weights

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        # Nothing slow ever sits between two collectives below (the CPU legs run after destroy_process_group); the explicit
        # timeout only bounds a genuinely wedged rank.
        dist.init_process_group("nccl", device_id=torch.device(dev), timeout=datetime.timedelta(minutes=20))

    # ---- engine with random-init weights of the CosyVoice2-0.5B-EU architecture ----
    flow, hift = B200Flow(dev), B200HiFT(dev)
    flow.load_state_dict(weights.to_torch(weights.make_flow_state()))
    hift.load_state_dict(weights.to_torch(weights.make_hift_state()))
    t2w = B200Token2Wav(flow, hift)
    eng = flow.eng

    # ---- this rank's shard: BATCH utterances, length-sorted (bucketing keeps neighbours similar) ----
    n_tokens = sorted(workload(rank, args.batch))
    utts = [weights.make_utterance(n, N_PROMPT, seed=rank * 100000 + i) for i, n in enumerate(n_tokens)]
    tokens = [torch.from_numpy(u["token"][0]) for u in utts]
    ptoks = [torch.from_numpy(u["prompt_token"][0]) for u in utts]
    pfeats = [torch.from_numpy(u["prompt_feat"][0]) for u in utts]
    embs = [torch.from_numpy(u["embedding"][0]) for u in utts]
    audio_s = sum(2 * n * 480 for n in n_tokens) / 24000.0
    B = len(n_tokens)
    max_total = max(n_tokens) + N_PROMPT
    mel_T = 2 * max(n_tokens)
    # the gather needs ONE shape on all ranks (ragged counts are undefined behaviour in NCCL): the longest utterance over all
    # ranks, which every rank computes for itself (workload() is deterministic)
    gmax_samples = 960 * max(max(workload(r, args.batch)) for r in range(world))

    # device-resident inputs for the kernel-only measurement
    tok_d = torch.zeros(B, max(n_tokens), dtype=torch.int32)
    for b, t in enumerate(tokens):
        tok_d[b, :t.numel()] = t
    tok_d = tok_d.to(dev)
    ptk_d = torch.stack(ptoks).to(torch.int32).to(dev)
    pf_d = torch.stack(pfeats).to(dev)
    emb_d = torch.stack(embs).to(dev)
    tl_d = torch.tensor(n_tokens, dtype=torch.int32, device=dev)
    pl_d = torch.full((B,), N_PROMPT, dtype=torch.int32, device=dev)
    fl_d = torch.full((B,), 2 * N_PROMPT, dtype=torch.int32, device=dev)
    mel_lens_d = (tl_d * 2).contiguous()

    def step_resident():
        (mel,) = flow._forward_device(tok_d, tl_d, ptk_d, pl_d, pf_d, fl_d, emb_d, B, max_total, mel_T, False, True)
        speech, _ = hift.inference(mel, lens=mel_lens_d)
        return speech, flow.last_launches + hift.last_launches

    gather_buf = {}

    def step_e2e():
        speech, lens = t2w.token2wav_batch(tokens, ptoks, pfeats, embs)
        if world > 1:                                        # final gather of (lengths, padded audio) on rank 0 (the only collective)
            if not gather_buf:                               # send / receive buffers are allocated once, not per step
                gather_buf["send"] = torch.zeros(B, gmax_samples, dtype=torch.float32, device=dev)
                gather_buf["lens_send"] = torch.zeros(B, dtype=torch.int32, device=dev)
                if rank == 0:
                    gather_buf["lens"] = [torch.empty_like(gather_buf["lens_send"]) for _ in range(world)]
                    gather_buf["aud"] = [torch.empty_like(gather_buf["send"]) for _ in range(world)]
            gather_buf["lens_send"].copy_(lens, non_blocking=True)
            gather_buf["send"][:, :speech.shape[1]].copy_(speech)
            dist.gather(gather_buf["lens_send"], gather_buf.get("lens"), dst=0)
            dist.gather(gather_buf["send"], gather_buf.get("aud"), dst=0)
        if "host" not in gather_buf or gather_buf["host"].shape != speech.shape:
            gather_buf["host"] = torch.empty(speech.shape, dtype=speech.dtype).pin_memory()   # pinned once, reused every step
        out = gather_buf["host"]
        out.copy_(speech, non_blocking=True)                 # D2H of the step's result into pinned memory
        torch.cuda.current_stream().synchronize()            # (also drains the gather on rank 0)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_per_step = 0
    for _ in range(max(args.warmup, 1)):
        _, launches_per_step = step_resident()
    barrier()

    # ---- timed region 1: device-resident inputs, CUDA events around exactly K steps -> `value` ----
    stop, clk = threading.Event(), []
    th = threading.Thread(target=sample_clocks, args=(stop, clk), daemon=True)
    th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # ---- the same K steps again with a CUDA-event pair around every launch (~7 k event records per step, a few ms of
    #      overhead, which is why it is not the `value` region): per-kernel-family device time for the roofline ----
    eng.lib.cv2_engine_set_profiling(eng.h, 1)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms_prof = e0.elapsed_time(e1)
    raw_ms = (C.c_double * NFAM)()
    raw_n = (C.c_longlong * NFAM)()
    eng.lib.cv2_engine_read_profile(eng.h, raw_ms, raw_n, NFAM)
    eng.lib.cv2_engine_set_profiling(eng.h, 0)
    fam_ms, fam_n = list(raw_ms[:len(FAMILIES)]), list(raw_n[:len(FAMILIES)])
    fam_ms[2] += sum(raw_ms[len(FAMILIES):])
    fam_n[2] += sum(raw_n[len(FAMILIES):])
    g256 = {G256[i]: {"ms_per_step": raw_ms[len(FAMILIES) + i] / args.steps, "launches_per_step": raw_n[len(FAMILIES) + i] / args.steps}
            for i in range(len(G256)) if raw_n[len(FAMILIES) + i] > 0}

    # ---- timed region 2: end to end through the public API with host buffers ----
    for _ in range(2):
        step_e2e()
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step_e2e()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    stop.set()
    h2d = flow.last_h2d_bytes
    d2h = out.numel() * 4 + 4 * B
    workspace_gb = {}
    for (kind, _), ent in eng.ws.items():
        workspace_gb[kind] = max(workspace_gb.get(kind, 0.0), ent[0].numel() / 1e9)

    # ---- BASELINE configs[4]: a 4096-utterance corpus sharded by utterance over the ranks (strong scaling): cost-balanced
    #      shards, length-bucketed batches of <= 64 through the public batched API (host inputs), audio packed into one flat
    #      buffer per rank, ONE NCCL gather to rank 0, D2H of everything there ----
    corpus = None
    if args.corpus > 0:
        crng = np.random.Generator(np.random.Philox(key=4096))
        c_tok = [int(round(25 * d)) for d in crng.uniform(4.0, 20.0, size=args.corpus)]
        plan, flat_len = shard.corpus_plan(c_tok, [N_PROMPT] * len(c_tok), world)
        mine = plan[rank]
        c_utts = {i: weights.make_utterance(c_tok[i], N_PROMPT, seed=7_000_000 + i) for i in mine["indices"]}

        def corpus_batch(idx):
            us = [c_utts[i] for i in idx]
            return t2w.token2wav_batch([torch.from_numpy(u["token"][0]) for u in us], [torch.from_numpy(u["prompt_token"][0]) for u in us],
                                       [torch.from_numpy(u["prompt_feat"][0]) for u in us], [torch.from_numpy(u["embedding"][0]) for u in us])

        flat = torch.zeros(flat_len, dtype=torch.float32, device=dev)
        recv = [torch.empty_like(flat) for _ in range(world)] if (rank == 0 and world > 1) else None
        host = torch.empty((world if rank == 0 else 0) * flat_len, dtype=torch.float32).pin_memory() if rank == 0 else None
        # warm-up: the largest batch (grows the workspaces to their final size) and the smallest one
        big = max(mine["batches"], key=lambda b: len(b) * (c_tok[b[-1]] + N_PROMPT) ** 2)
        corpus_batch(big)
        corpus_batch(mine["batches"][0])
        barrier()
        e0.record()
        shard.run_corpus_shard(mine, c_tok, corpus_batch, flat)
        if world > 1:
            shard.gather_flat(flat, dst=0, out=recv)
        if rank == 0:
            for r in range(world):
                host[r * flat_len:(r + 1) * flat_len].copy_(recv[r] if world > 1 else flat, non_blocking=True)
        e1.record()
        barrier()
        ms_corpus = e0.elapsed_time(e1)
        corpus = {"ms": ms_corpus, "utterances": len(c_tok), "audio_s": sum(c_tok) * 960 / 24000.0,
                  "batches_this_rank": len(mine["batches"]), "utterances_this_rank": len(mine["indices"]),
                  "gather_bytes": int(4 * flat_len * (world - 1)) if world > 1 else 0, "d2h_bytes": int(4 * flat_len * world)}
        del flat, recv, host, c_utts
        torch.cuda.empty_cache()

    # ---- reduce over ranks: the last collectives; after them ranks >= 1 are done ----
    red = torch.tensor([ms, ms_e2e, corpus["ms"] if corpus else 0.0], dtype=torch.float64, device=dev)
    aud = torch.tensor([audio_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(aud, op=dist.ReduceOp.SUM)
    ms, ms_e2e = float(red[0]), float(red[1])
    if corpus:
        corpus["ms"] = float(red[2])
    total_audio = float(aud[0])
    if world > 1:
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if rank != 0:
        return          # every other rank exits here: rank 0's side legs and CPU baseline run with the box to themselves

    peaks, peak_src = load_peaks()
    side = {}
    if not args.no_side_legs:
        side = side_legs(args, t2w, flow, hift, dev, peaks, clk, chunk_schedule, weights)

    flops = algorithmic_flops(n_tokens)
    fam = {n: {"ms_per_step": fam_ms[i] / args.steps, "launches_per_step": fam_n[i] / args.steps,
               "algorithmic_tflop_per_step": flops[n] / 1e12,
               "tflops": (flops[n] / 1e12) / (fam_ms[i] / args.steps / 1e3) if fam_ms[i] > 0 else None}
           for i, n in enumerate(FAMILIES) if fam_n[i] > 0}
    # dominant KERNEL by device time: gemm_tap<256> is one template with many epilogue specialisations that serve different
    # layers; its two big instances (QKV projection, out-proj + residual + LayerNorm) are ranked on their own
    rows2 = 20.0 * sum(2 * (n + N_PROMPT) for n in n_tokens)          # estimator rows x Euler steps x CFG
    spec_flops = {"qkv_split": rows2 * 56 * 2 * 256 * 1536.0, "res+ln_emit": rows2 * 56 * 2 * 512 * 256.0}
    cand = {n: v for n, v in fam.items() if n != "gemm_tap<256>"}
    for sname, fl in spec_flops.items():
        if sname in g256:
            cand["gemm_tap<256>:" + sname] = {"ms_per_step": g256[sname]["ms_per_step"], "launches_per_step": g256[sname]["launches_per_step"],
                                              "algorithmic_tflop_per_step": fl / 1e12,
                                              "tflops": fl / 1e12 / (g256[sname]["ms_per_step"] / 1e3)}
    dom = max(cand, key=lambda n: cand[n]["ms_per_step"])
    d = cand[dom]
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    ach = d["tflops"] or 0.0
    traffic = kernel_traffic().get(dom, {})
    roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": traffic.get("dram_bytes_per_launch"), "traffic_source": traffic.get("source"),
                "peak_source": peak_src + " sustained dense bf16 (kernel timed inside a long step)",
                "avg_launch_ms": d["ms_per_step"] / d["launches_per_step"],
                "algorithmic_gflop_per_launch": 1e3 * d["algorithmic_tflop_per_step"] / d["launches_per_step"],
                "share_of_step": d["ms_per_step"] / (ms_prof / args.steps), "ms_per_step_profiled_pass": ms_prof / args.steps}
    # bandwidth-bound vocoder kernels: algorithmic bytes (fp32 I/O per output sample) / event-timed launch
    samples = float(sum(2 * n * 480 for n in n_tokens))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_bytes = {"source_stft": 22.0, "istft": 22.0, "nsf_source": 4.0 + 4.0 / 480}
    hbm_kernels = {}
    for k, bps in hbm_bytes.items():
        if k in fam and fam[k]["ms_per_step"] > 0:
            gbs = samples * bps / (fam[k]["ms_per_step"] / fam[k]["launches_per_step"] * 1e-3) / 1e9
            hbm_kernels[k] = {"algorithmic_bytes_per_sample": bps, "ms_per_launch": fam[k]["ms_per_step"] / fam[k]["launches_per_step"],
                              "achieved_gbs": gbs, "peak_gbs": hbm_peak, "frac": gbs / hbm_peak}
    if "one_directional" in side:
        od = side["one_directional"]
        hbm_kernels["one_directional_peaks_gbs_measured_here"] = od
        if "source_stft" in hbm_kernels and "write_only_fill" in od:
            hbm_kernels["source_stft"]["frac_of_write_only"] = hbm_kernels["source_stft"]["achieved_gbs"] / od["write_only_fill"]
        if "istft" in hbm_kernels and "read_only_sum" in od:
            hbm_kernels["istft"]["frac_of_read_only"] = hbm_kernels["istft"]["achieved_gbs"] / od["read_only_sum"]
    hbm_kernels["note"] = ("peak = measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs); nsf_source in production mode draws its "
                           "noise in-kernel (9 sines + 9 normals per 4-byte sample), i.e. it is ALU/MUFU bound there and only "
                           "bandwidth bound in parity mode (36 B/sample noise read)")
    total_flop = sum(flops.values()) * world
    cfg = workload_config(world, args.batch)
    cfg["utterances_per_gpu"], cfg["audio_seconds_per_step_per_gpu"] = B, audio_s
    line = {
        "metric": "token2wav_audio_seconds_per_second", "value": total_audio * args.steps / (ms / 1e3), "unit": "audio-s/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16 operands, fp32 accumulate (tcgen05 kind::f16)",
        "data": "synthetic", "config": cfg,
        "e2e": {"value": total_audio * args.steps / (ms_e2e / 1e3), "unit": "audio-s/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline,
        "families": fam,
        "hbm_kernels": hbm_kernels,
        "gemm256_by_epilogue": g256,
        "whole_step_tflops": total_flop / 1e12 / (ms / args.steps / 1e3) / world,
        "whole_step_frac_of_peak": total_flop / 1e12 / (ms / args.steps / 1e3) / world / peak,
        "workspace_gb": workspace_gb,
        "clocks": clocks_summary(clk),
    }
    if corpus:
        line["corpus4096"] = {
            "workload": f"BASELINE configs[4]: {corpus['utterances']} utterances, U[4,20] s + 75-token prompt, sharded by cost over "
                        f"{world} GPU(s) (shard_by_cost), length-bucketed batches <= 64 (bucket_batches), host inputs, one NCCL "
                        "gather of the packed audio to rank 0, D2H there; ONE timed pass, strong scaling",
            "scaling": "strong", "audio_s_per_s": corpus["audio_s"] / (corpus["ms"] / 1e3), "seconds": corpus["ms"] / 1e3,
            "audio_seconds": corpus["audio_s"], "utterances_rank0": corpus["utterances_this_rank"],
            "batches_rank0": corpus["batches_this_rank"], "gather_bytes": corpus["gather_bytes"], "d2h_bytes": corpus["d2h_bytes"]}
    for k in ("rtf_batch1", "stream32", "prompt_mel"):
        if k in side:
            line[k] = side[k]
    if not args.no_cpu_baseline:
        # the GPU legs are over and the other ranks have exited: all host cores belong to this leg
        cores = os.cpu_count()
        torch.set_num_threads(cores)        # torchrun exports OMP_NUM_THREADS=1
        runs = [cpu_reference_run(200, cores) for _ in range(4)]       # 1 warm-up + 3 timed
        dts = sorted(r[0] for r in runs[1:])
        dt, a = dts[1], runs[1][1]
        line["cpu_baseline"] = {"value": a / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
                                "sample": "one 8 s utterance (200 tokens + 75-token prompt, configs[0]) through the oracle port of "
                                          "the reference CPU token2wav (torch fp32, all host threads): 1 warm-up, median of 3",
                                "seconds": dt, "seconds_all": [r[0] for r in runs]}
    print(json.dumps(line))


def side_legs(args, t2w, flow, hift, dev, peaks, clk, chunk_schedule, weights):
    """Rank 0 only, after every collective: configs[1] latency, configs[3] streaming, prompt features, one-directional HBM peaks."""
    import torch
    out = {}
    peak_burst = float(peaks.get("bf16_tflops", 1590.0))
    # ---- configs[1]: batch-1 latency / RTF (single 10 s utterance) ----
    u1 = weights.make_utterance(250, N_PROMPT, seed=99)
    a1 = [torch.from_numpy(u1[k][0]) for k in ("token", "prompt_token", "prompt_feat", "embedding")]
    from cosyvoice2_eu_b200 import GraphedToken2Wav
    g1 = GraphedToken2Wav(t2w)                     # CUDA-graph replay of the whole token2wav (host -> wav on host)
    for _ in range(3):
        g1([a1[0]], [a1[1]], [a1[2]], [a1[3]])
    torch.cuda.synchronize()
    lat = []
    for _ in range(7):
        t0 = time.perf_counter()
        w, _ = g1([a1[0]], [a1[1]], [a1[2]], [a1[3]])
        w.cpu()
        lat.append(time.perf_counter() - t0)
    lat1 = float(np.median(lat))
    lat_eager = []
    for _ in range(3):
        t0 = time.perf_counter()
        w, _ = t2w.token2wav_batch([a1[0]], [a1[1]], [a1[2]], [a1[3]])
        w.cpu()
        lat_eager.append(time.perf_counter() - t0)
    fl1 = sum(algorithmic_flops([250]).values())
    out["rtf_batch1"] = {"workload": "BASELINE configs[1]: single 10 s utterance (250 tokens + 75-token prompt), batch 1",
                         "latency_s": lat1, "rtf": lat1 / 10.0, "audio_s_per_s": 10.0 / lat1,
                         "how": "CUDA-graph replay, host in / host out",
                         "latency_s_eager_launches": float(np.median(lat_eager)),
                         "roofline": {"bound": "tensor", "algorithmic_tflop": fl1 / 1e12, "achieved": fl1 / 1e12 / lat1,
                                      "peak": peak_burst, "unit": "TFLOP/s", "frac": fl1 / 1e12 / lat1 / peak_burst,
                                      "note": "whole call, host to host; burst peak (a short call is not power capped); 650 x 2 "
                                              "estimator rows are 12 row tiles of 128 for 148 SMs: occupancy bound, not pipe bound"}}

    # ---- configs[3]: 32 concurrent streaming sessions (hop 25, lookahead 3, mel cache 8, source cache 3840), N = 250 ----
    n_sess, n_tok_s = 32, 250
    sess = [weights.make_utterance(n_tok_s, N_PROMPT, seed=5000 + i) for i in range(n_sess)]
    sess = [{k: torch.from_numpy(v) for k, v in u.items()} for u in sess]
    sched = chunk_schedule(n_tok_s, N_PROMPT)

    def run_streams(group):
        for i in range(n_sess):
            t2w.hift_cache_dict[f"bench{i}"] = None
        lat, total = [], 0
        for (n_vis, off, fin) in sched:
            t0 = time.perf_counter()
            reqs = [dict(token=u["token"][:, :n_vis], prompt_token=u["prompt_token"], prompt_feat=u["prompt_feat"],
                         embedding=u["embedding"], token_offset=off, uuid=f"bench{i}") for i, u in enumerate(sess)]
            outs = t2w.token2wav_stream_batch(reqs, finalize=fin, group=group)
            host = [o.cpu() for o in outs]
            lat.append(time.perf_counter() - t0)
            total += sum(o.shape[1] for o in host)
        return lat, total

    def timed_streams(group):
        run_streams(group)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lat, total = run_streams(group)
        return lat, total, time.perf_counter() - t0
    # the reference's schedule as it stands (every chunk recomputes the whole prefix), then the same chunks computed
    # incrementally on a StreamGroup (k / v^T + causal-conv state of earlier frames stay on the device, row F1)
    lat_p, total_p, t_prefix = timed_streams(None)
    T_need = 2 * (N_PROMPT + max(nv for nv, _, fin in sched if not fin) - 3) if len(sched) > 1 else 128
    group = flow.open_stream_group(n_sess, max_mel_frames=T_need)
    lat_s, total_samples, t_stream = timed_streams(group)
    state_gb = group.state.numel() / 1e9
    del group
    torch.cuda.empty_cache()
    # algorithmic FLOPs of a session in the INCREMENTAL unit: every frame of the non-final chunks computed once
    # with the block-causal attention term, plus the full final pass
    def est_flops(T, causal):
        att = 114688.0 * ((T + 50) / 2 if causal else T)
        return 20.0 * T * (132161536.0 + att)
    T_last_nonfinal = 2 * (N_PROMPT + sched[-2][0] - 3) if len(sched) > 1 else 0
    T_final = 2 * (N_PROMPT + n_tok_s)
    fl_sess = est_flops(T_last_nonfinal, True) + est_flops(T_final, False) + 612.3e6 * 2 * n_tok_s
    out["stream32"] = {"workload": "BASELINE configs[3]: 32 concurrent sessions x 250 tokens (10 s), hop 25, chunk schedule of "
                                   "CosyVoice2Model.tts, every step = one ragged batch over all sessions, chunks copied to the host; "
                                   "non-final chunks incremental on a StreamGroup, final chunk = the reference's full-attention pass",
                       "audio_s_per_s": total_samples / 24000.0 / t_stream, "chunks": len(sched),
                       "chunk_latency_s_median": float(np.median(lat_s)), "first_chunk_latency_s": lat_s[0], "wall_s": t_stream,
                       "chunk_latency_s": [round(x, 5) for x in lat_s], "stream_state_gb": state_gb,
                       "prefix_recompute": {"audio_s_per_s": total_p / 24000.0 / t_prefix, "wall_s": t_prefix,
                                            "chunk_latency_s_median": float(np.median(lat_p)), "first_chunk_latency_s": lat_p[0],
                                            "note": "same sessions with every chunk recomputing the prefix (the reference's schedule "
                                                    "as it stands, model.py:351-381)"},
                       "roofline": {"bound": "tensor", "unit": "TFLOP/s", "algorithmic_tflop": n_sess * fl_sess / 1e12,
                                    "achieved": n_sess * fl_sess / 1e12 / t_stream, "peak": peak_burst,
                                    "frac": n_sess * fl_sess / 1e12 / t_stream / peak_burst,
                                    "note": "incremental unit: each non-final frame counted once (block-causal attention) + the "
                                            "full final pass + the vocoder; prefix recomputation is not counted as useful work"}}
    # one-directional HBM peaks (the copy figure is a read+write mix; the STFT is 82 % writes and the iSTFT 82 % reads)
    try:
        buf = torch.empty(1 << 28, dtype=torch.float32, device=dev)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wr, rd = [], []
        for _ in range(5):
            ev0.record(); buf.fill_(1.0); ev1.record(); torch.cuda.synchronize(); wr.append(buf.numel() * 4 / ev0.elapsed_time(ev1) / 1e6)
            ev0.record(); buf.sum(); ev1.record(); torch.cuda.synchronize(); rd.append(buf.numel() * 4 / ev0.elapsed_time(ev1) / 1e6)
        del buf
        out["one_directional"] = {"write_only_fill": max(wr), "read_only_sum": max(rd)}
    except Exception as exc:   # never let a side measurement break the bench line
        out["one_directional"] = {"error": str(exc)}
    # SURVEY.md 8f row F2 (the step before the path): prompt log-mel of 64 prompts of U[3, 30] s
    try:
        out["prompt_mel"] = bench_prompt_mel(dev, peaks, clk, with_cpu=not args.no_cpu_baseline)
    except Exception as exc:
        out["prompt_mel"] = {"error": str(exc)}
    return out


def bench_prompt_mel(dev, peaks, clk, with_cpu=True, reps=10):
    import torch
    from cosyvoice2_eu_b200 import extract_speech_feat_batch, frontend
    rng = np.random.Generator(np.random.Philox(key=4242))
    lens = [int(24000 * d) for d in rng.uniform(3.0, 30.0, size=63)] + [24000 * 30]
    g = torch.Generator().manual_seed(5)
    waves = [(torch.rand(n, generator=g) * 2 - 1) * 0.5 for n in lens]
    max_len = max(lens)
    wav = torch.zeros(len(lens), max_len, dtype=torch.float32, device=dev)
    for i, w in enumerate(waves):
        wav[i, :lens[i]] = w.to(dev)
    n = torch.tensor(lens, dtype=torch.int32, device=dev)
    for _ in range(3):
        mel, mel_len = frontend._run(wav, n, max_len)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        mel, mel_len = frontend._run(wav, n, max_len)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    frames = int(mel_len.sum())
    t0 = time.perf_counter()
    for _ in range(reps):
        m2, l2 = extract_speech_feat_batch(waves, device=dev)
        l2.cpu()
    ms_host = (time.perf_counter() - t0) / reps * 1e3
    audio = sum(lens) / 24000.0
    flop = frames * 2.0 * 2 * 961 * 961                 # folded real DFT: 2 x 961 x 961 FMAs per frame
    sm_mhz = clocks_summary(clk).get("sm_mhz") or 1850.0
    fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12     # 128 FP32 lanes per SM, FMA = 2 flop, at the clock sampled under load
    out = {"workload": "64 prompts, U[3,30] s at 24 kHz, one launch pair (pm_dft_mag + pm_mel_log)", "frames": frames,
           "ms_per_call": ms, "audio_s_per_s": audio / (ms / 1e3), "fp32_tflops_algorithmic": flop / 1e12 / (ms / 1e3),
           "fp32_peak_tflops": fp32_peak, "frac": flop / 1e12 / (ms / 1e3) / fp32_peak, "bound": "fp32 FMA",
           "e2e_host_buffers": {"ms_per_call": ms_host, "audio_s_per_s": audio / (ms_host / 1e3), "h2d_bytes": int(len(lens) * max_len * 4),
                                "d2h_bytes": 4 * len(lens)}}
    # the 16 -> 24 kHz resampler in front of it (HBM bound: 4 B read per input sample + 4 B written per output sample)
    from cosyvoice2_eu_b200 import resample_16k_to_24k
    lens16 = [n * 2 // 3 for n in lens]
    x16 = torch.rand(len(lens16), max(lens16), device=dev) - 0.5
    n16 = torch.tensor(lens16, dtype=torch.int32, device=dev)
    for _ in range(3):
        resample_16k_to_24k(x16, device=dev, lengths=n16)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        resample_16k_to_24k(x16, device=dev, lengths=n16)
    e1.record()
    torch.cuda.synchronize()
    ms_rs = e0.elapsed_time(e1) / reps
    rs_bytes = 4.0 * x16.numel() * 2.5          # padded rows are read and written in full
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    out["resample_16k_24k"] = {"ms_per_call": ms_rs, "achieved_gbs": rs_bytes / (ms_rs * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                               "frac": rs_bytes / (ms_rs * 1e-3) / 1e9 / hbm_peak, "bound": "hbm",
                               "note": "includes the torch.empty of the output; 10 B per input sample"}
    if with_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import prompt_mel_oracle as PO     # checker timed as the CPU baseline (bench.py's cpu_baseline leg)
        w0 = waves[-1].numpy()
        t0 = time.perf_counter()
        ref = PO.mel_spectrogram(w0)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": 30.0 / dt, "unit": "audio-s/s", "cores": 1, "kind": "port", "sample": "one 30 s prompt",
                               "max_abs_logmel_diff_vs_gpu": float(np.abs(mel[-1, :ref.shape[2]].cpu().numpy() - ref[0].T).max())}
    return out


if __name__ == "__main__":
    main()
