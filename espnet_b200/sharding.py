"""Utterance sharding across ranks and the single exchange of the path: an all-gather of fixed-width hypothesis records.

The reference parallelises decoding only by splitting the key file over OS processes (egs2/TEMPLATE/asr1/asr.sh:1591-1618)
and concatenating result files; here rank r decodes utterances r, r+world, ... and the n-best lists are gathered once with
torch.distributed (NCCL on GPUs, gloo in the CPU tests).  Records are int32 [n_local_padded, nbest, 2 + max_tokens]:
(n_tokens, float32 score bits, tokens..., -1 padding).  Unused n-best entries have n_tokens = -1; padding ROWS (ranks holding fewer
utterances) have n_tokens = -2 in entry 0, so that a real utterance with an empty n-best list (nothing ended, beam_search.py:467-471)
keeps its row and later utterances are not shifted.
"""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


PAD_ROW = -2   # n_tokens marker of a padding row (entry 0)


def shard_indices(n_utts: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_utts, world))


def pack_hypotheses(nbest_lists: Sequence[Sequence[Tuple[List[int], float]]], nbest: int, max_tokens: int, rows: int = None,
                    device="cpu") -> torch.Tensor:
    rows = len(nbest_lists) if rows is None else rows
    rec = torch.full((rows, nbest, 2 + max_tokens), -1, dtype=torch.int32)
    rec[len(nbest_lists):, 0, 0] = PAD_ROW
    for i, hyps in enumerate(nbest_lists):
        for j, (toks, score) in enumerate(list(hyps)[:nbest]):
            toks = list(toks)[:max_tokens]
            rec[i, j, 0] = len(toks)
            rec[i, j, 1] = torch.tensor(score, dtype=torch.float32).view(torch.int32)
            if toks:
                rec[i, j, 2:2 + len(toks)] = torch.tensor(toks, dtype=torch.int32)
    return rec.to(device)


def unpack_hypotheses(rec: torch.Tensor, nbest: int):
    rec = rec.cpu()
    out = []
    for i in range(rec.shape[0]):
        hyps = []
        for j in range(nbest):
            n = int(rec[i, j, 0])
            if n < 0:
                continue
            score = float(rec[i, j, 1].view(torch.float32))
            hyps.append((rec[i, j, 2:2 + n].tolist(), score))
        if int(rec[i, 0, 0]) != PAD_ROW:
            out.append(hyps)
    return out


def gather_hypotheses(rec: torch.Tensor, world: int) -> torch.Tensor:
    """All ranks must pass records of the same shape (pad rows with pack_hypotheses(rows=...)). Returns [world, rows, nbest, width]."""
    if world == 1:
        return rec.unsqueeze(0)
    # equalise the row count (ranks may hold one utterance fewer)
    n = torch.tensor([rec.shape[0]], dtype=torch.int64, device=rec.device)
    dist.all_reduce(n, op=dist.ReduceOp.MAX)
    rows = int(n.item())
    if rec.shape[0] < rows:
        pad = torch.full((rows - rec.shape[0],) + tuple(rec.shape[1:]), -1, dtype=rec.dtype, device=rec.device)
        pad[:, 0, 0] = PAD_ROW
        rec = torch.cat([rec, pad], 0)
    out = torch.empty((world,) + tuple(rec.shape), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out.view(world * rows, *rec.shape[1:]), rec.contiguous())
    return out
