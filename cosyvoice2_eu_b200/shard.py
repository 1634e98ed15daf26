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
