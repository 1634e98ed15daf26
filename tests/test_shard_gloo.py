"""CPU, world_size 2 over gloo: the multi-GPU host logic (cost-balanced utterance sharding + the final gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, n_tokens, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cosyvoice2_eu_b200 import shard
    shards = shard.shard_by_cost(n_tokens, [75] * len(n_tokens), world)
    mine = shards[rank]
    # fake "audio": utterance i -> ramp of length 960*n_tokens[i] scaled by i
    L = max(960 * n_tokens[i] for i in mine)
    speech = torch.zeros(len(mine), L)
    for k, i in enumerate(mine):
        speech[k, :960 * n_tokens[i]] = float(i + 1)
    lens = torch.tensor([960 * n_tokens[i] for i in mine], dtype=torch.int32)
    s, l = shard.gather_audio(speech, lens, dst=0)
    if rank == 0:
        ok = True
        seen = []
        for r in range(world):
            for k, i in enumerate(shards[r]):
                n = int(l[r][k])
                ok &= n == 960 * n_tokens[i]
                ok &= bool((s[r][k, :n] == float(i + 1)).all()) and bool((s[r][k, n:] == 0).all())
                seen.append(i)
        ok &= sorted(seen) == list(range(len(n_tokens)))
        torch.save(ok, tmp)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2(tmp_path):
    n_tokens = [100, 500, 250, 333, 120, 480, 199, 410, 275]
    tmp = str(tmp_path / "ok.pt")
    mp.spawn(_worker, args=(2, 29531, n_tokens, tmp), nprocs=2, join=True)
    assert torch.load(tmp) is True


def _corpus_worker(rank, world, port, n_tokens, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cosyvoice2_eu_b200 import shard
    plan, flat_len = shard.corpus_plan(n_tokens, [75] * len(n_tokens), world, max_batch=4)

    def batch_fn(idx):       # stands in for B200Token2Wav.token2wav_batch: utterance i -> constant i + 1, zero padded
        L = max(960 * n_tokens[i] for i in idx)
        sp = torch.zeros(len(idx), L)
        for k, i in enumerate(idx):
            sp[k, :960 * n_tokens[i]] = float(i + 1)
        return sp, None

    flat = torch.zeros(flat_len)
    shard.run_corpus_shard(plan[rank], n_tokens, batch_fn, flat)
    got = shard.gather_flat(flat, dst=0)
    if rank == 0:
        audio = shard.unpack_corpus(got, plan, n_tokens)
        ok = sorted(audio) == list(range(len(n_tokens)))
        for i, a in audio.items():
            ok &= a.numel() == 960 * n_tokens[i] and bool((a == float(i + 1)).all())
        ok &= all(g.numel() == flat_len for g in got)
        torch.save(ok, tmp)
    dist.barrier()
    dist.destroy_process_group()


def test_corpus_plan_run_and_flat_gather_world2(tmp_path):
    """configs[4] host logic at world_size 2 (gloo): plan -> per-rank bucketed batches -> packed flat audio -> one gather with
    the SAME shape on every rank -> every utterance recovered on rank 0."""
    n_tokens = [100, 500, 250, 333, 120, 480, 199, 410, 275, 101, 455, 222, 318]
    tmp = str(tmp_path / "ok2.pt")
    mp.spawn(_corpus_worker, args=(2, 29537, n_tokens, tmp), nprocs=2, join=True)
    assert torch.load(tmp) is True


def test_cost_balance_and_buckets():
    from cosyvoice2_eu_b200 import shard
    import numpy as np
    rng = np.random.default_rng(0)
    n = [int(round(25 * d)) for d in rng.uniform(4, 20, 4096)]
    for world in (2, 4, 8):
        sh = shard.shard_by_cost(n, [75] * len(n), world)
        assert sorted(i for s in sh for i in s) == list(range(len(n)))
        loads = [sum(shard.utterance_cost(n[i], 75) for i in s) for s in sh]
        assert max(loads) / min(loads) < 1.001
        for s in sh:
            assert all(n[a] <= n[b] for a, b in zip(s, s[1:]))
            for b in shard.bucket_batches(s, n):
                assert len(b) <= 64 and n[b[-1]] <= 1.35 * n[b[0]]
