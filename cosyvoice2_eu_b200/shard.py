"""Utterance sharding for multi-GPU token2wav: utterances are independent, weights are replicated, so a corpus
is partitioned by utterance with no data-path collective; the only exchange is the final gather of (lengths, padded audio).
Host logic only -- works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist

# cost model from the algorithmic FLOPs of SURVEY.md 8(d): estimator linear + attention terms dominate
_A, _B = 2 * 10 * 132.16e6, 2 * 10 * 114688.0


def utterance_cost(n_tokens, n_prompt):
    T = 2 * (n_tokens + n_prompt)
    return _A * T + _B * T * T + 612.3e6 * 2 * n_tokens


def shard_by_cost(n_tokens, n_prompts, world):
    """Greedy longest-processing-time assignment: equal FLOPs per rank, not equal counts.
    Returns a list (per rank) of utterance indices, each sorted by length (length bucketing)."""
    order = sorted(range(len(n_tokens)), key=lambda i: -utterance_cost(n_tokens[i], n_prompts[i]))
    loads = [0.0] * world
    shards = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: loads[k])
        shards[r].append(i)
        loads[r] += utterance_cost(n_tokens[i], n_prompts[i])
    return [sorted(s, key=lambda i: n_tokens[i]) for s in shards]


def bucket_batches(indices, n_tokens, max_batch=64, max_ratio=1.35):
    """Split a length-sorted shard into batches whose longest/shortest length ratio stays below max_ratio."""
    out, cur = [], []
    for i in indices:
        if cur and (len(cur) >= max_batch or n_tokens[i] > max_ratio * n_tokens[cur[0]]):
            out.append(cur)
            cur = []
        cur.append(i)
    if cur:
        out.append(cur)
    return out


def gather_audio(speech, lengths, dst=0, group=None):
    """speech [n_local, L] (zero padded), lengths int32 [n_local] -> on dst: (list of per-rank speech, list of per-rank lengths).
    One gather of the sizes, one of the lengths, one of the audio."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = speech.device
    meta = torch.tensor([speech.shape[0], speech.shape[1]], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    n_max = int(max(m[0] for m in metas))
    l_max = int(max(m[1] for m in metas))
    pad = torch.zeros(n_max, l_max, dtype=speech.dtype, device=dev)
    pad[:speech.shape[0], :speech.shape[1]] = speech
    lpad = torch.zeros(n_max, dtype=torch.int32, device=dev)
    lpad[:lengths.numel()] = lengths.to(dev, torch.int32)
    outs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    louts = [torch.empty_like(lpad) for _ in range(world)] if rank == dst else None
    dist.gather(lpad, louts, dst=dst, group=group)
    dist.gather(pad, outs, dst=dst, group=group)
    if rank != dst:
        return None, None
    res_s, res_l = [], []
    for r in range(world):
        n = int(metas[r][0])
        res_s.append(outs[r][:n])
        res_l.append(louts[r][:n])
    return res_s, res_l


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE configs[4]: a corpus sharded by utterance over the ranks of one box, one gather at the end
# ---------------------------------------------------------------------------------------------------------------------
SAMPLES_PER_TOKEN = 2 * 480          # token_mel_ratio x hop: 960 samples (40 ms at 24 kHz) per speech token


def corpus_plan(n_tokens, n_prompts, world, max_batch=64, max_ratio=1.35):
    """The whole job as every rank computes it for itself (deterministic, no collective): per rank the utterance indices
    (cost balanced, length sorted), the length-bucketed batches, and where each utterance's samples go in that rank's flat
    audio buffer.  `flat_len` = the largest per-rank sample count: every rank allocates that much, so the final gather has one
    shape on all ranks (gathering ragged shapes is undefined behaviour in NCCL)."""
    shards = shard_by_cost(n_tokens, n_prompts, world)
    plan = []
    for r in range(world):
        offs, o = {}, 0
        for i in shards[r]:
            offs[i] = o
            o += SAMPLES_PER_TOKEN * n_tokens[i]
        plan.append({"indices": shards[r], "batches": bucket_batches(shards[r], n_tokens, max_batch, max_ratio), "offsets": offs,
                     "samples": o})
    flat_len = max(p["samples"] for p in plan)
    return plan, flat_len


def run_corpus_shard(plan_r, n_tokens, batch_fn, flat):
    """Run one rank's batches: batch_fn(indices) -> (speech [b, L] on the device, zero padded; anything else is ignored);
    every utterance's 960 * n_tokens samples are packed into `flat` at the planned offset."""
    for batch in plan_r["batches"]:
        speech = batch_fn(batch)[0]
        for k, i in enumerate(batch):
            n = SAMPLES_PER_TOKEN * n_tokens[i]
            o = plan_r["offsets"][i]
            flat[o:o + n].copy_(speech[k, :n], non_blocking=True)
    return flat


def gather_flat(flat, dst=0, group=None, out=None):
    """The one collective of the path: every rank's flat audio buffer (same length everywhere, see corpus_plan) to `dst`.
    `out`: optional preallocated receive buffers (list of world tensors like `flat`) reused across calls."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rank == dst and out is None:
        out = [torch.empty_like(flat) for _ in range(world)]
    dist.gather(flat, out if rank == dst else None, dst=dst, group=group)
    return out if rank == dst else None


def unpack_corpus(gathered, plan, n_tokens):
    """dst side: {utterance index: 1-D audio tensor (a view into the gathered buffers)}."""
    res = {}
    for r, p in enumerate(plan):
        for i in p["indices"]:
            o = p["offsets"][i]
            res[i] = gathered[r][o:o + SAMPLES_PER_TOKEN * n_tokens[i]]
    return res
