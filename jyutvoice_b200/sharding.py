"""Multi-GPU plumbing: utterances are independent, so the path shards with no data-path collective
(SURVEY.md section 8e).  One process per GPU; the only collective is the final gather of results."""
from typing import List, Sequence

import torch
import torch.distributed as dist


def cost(frames: int) -> float:
    """Relative work of one utterance: dense part ~T, attention ~T^2 (BASELINE.md section 4:
    132,161,536*T + 114,688*T^2 per estimator row)."""
    return frames * (1.0 + 114688.0 / 132161536.0 * frames)


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first assignment of utterance indices to ranks, balancing `cost`.
    Deterministic; every rank computes the same plan without communicating."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0.0] * world_size
    plan: List[List[int]] = [[] for _ in range(world_size)]
    cap = (len(lengths) + world_size - 1) // world_size  # equal batch sizes (+-1): gather buffers stay rectangular
    for i in order:
        r = min((k for k in range(world_size) if len(plan[k]) < cap), key=lambda k: (loads[k], k))
        plan[r].append(i)
        loads[r] += cost(int(lengths[i]))
    return [sorted(p) for p in plan]


def gather_waveforms(wav: torch.Tensor, wav_lengths: torch.Tensor, indices: Sequence[int], total: int, max_len: int):
    """All ranks contribute their shard ([n_r, L_r] padded waveforms, lengths, global utterance ids);
    every rank returns ([total, max_len] waveforms in the original order, [total] lengths).
    NCCL (or gloo on CPU) all_gather over equal-sized padded buffers."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    n_max = (total + world - 1) // world
    dev = wav.device
    buf = torch.zeros((n_max, max_len), dtype=wav.dtype, device=dev)
    lens = torch.zeros((n_max,), dtype=torch.int64, device=dev)
    ids = torch.full((n_max,), -1, dtype=torch.int64, device=dev)
    n = wav.shape[0]
    if n > n_max:
        raise ValueError("shard larger than the balanced maximum")
    buf[:n, : wav.shape[1]] = wav
    lens[:n] = wav_lengths.to(dev)
    ids[:n] = torch.as_tensor(list(indices), dtype=torch.int64, device=dev)
    if world > 1:
        bufs = [torch.empty_like(buf) for _ in range(world)]
        lenss = [torch.empty_like(lens) for _ in range(world)]
        idss = [torch.empty_like(ids) for _ in range(world)]
        dist.all_gather(bufs, buf)
        dist.all_gather(lenss, lens)
        dist.all_gather(idss, ids)
    else:
        bufs, lenss, idss = [buf], [lens], [ids]
    out = torch.zeros((total, max_len), dtype=wav.dtype, device=dev)
    out_len = torch.zeros((total,), dtype=torch.int64, device=dev)
    for b, l, i in zip(bufs, lenss, idss):
        keep = i >= 0
        out[i[keep]] = b[keep]
        out_len[i[keep]] = l[keep]
    return out, out_len
