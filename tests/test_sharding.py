"""CPU: the N>1 path (shard plan + result gather) with world_size-2 gloo."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jyutvoice_b200.sharding import shard_utterances, gather_waveforms


def test_plan_is_a_balanced_partition():
    lens = [300, 120, 290, 310, 50, 299, 301, 10, 500]
    for w in (1, 2, 4, 8):
        plan = shard_utterances(lens, w)
        assert sorted(i for p in plan for i in p) == list(range(len(lens)))
        assert max(len(p) for p in plan) - min(len(p) for p in plan) <= 2
    assert shard_utterances(lens, 2) == shard_utterances(lens, 2)


def _fake_vocoder(length, idx):
    return torch.arange(length * 4, dtype=torch.float32) * 0.001 + idx


def _worker(rank, world, port, lens, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = shard_utterances(lens, world)
    mine = plan[rank]
    max_len = max(lens) * 4
    wav = torch.zeros(len(mine), max_len)
    for j, i in enumerate(mine):
        wav[j, : lens[i] * 4] = _fake_vocoder(lens[i], i)
    wl = torch.tensor([lens[i] * 4 for i in mine])
    out, out_len = gather_waveforms(wav, wl, mine, len(lens), max_len)
    torch.save((out, out_len), os.path.join(out_dir, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gather_world2(tmp_path):
    lens = [7, 3, 9, 4, 8]
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, lens, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        out, out_len = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert out_len.tolist() == [l * 4 for l in lens]
        for i, l in enumerate(lens):
            assert torch.equal(out[i, : l * 4], _fake_vocoder(l, i))
            assert float(out[i, l * 4:].abs().sum()) == 0.0
