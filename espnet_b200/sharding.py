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
    import numpy as np

    rows = len(nbest_lists) if rows is None else rows
    rec = np.full((rows, nbest, 2 + max_tokens), -1, dtype=np.int32)
    rec[len(nbest_lists):, 0, 0] = PAD_ROW
    for i, hyps in enumerate(nbest_lists):
        for j, (toks, score) in enumerate(list(hyps)[:nbest]):
            toks = list(toks)[:max_tokens]
            rec[i, j, 0] = len(toks)
            rec[i, j, 1] = np.float32(score).view(np.int32)
            if toks:
                rec[i, j, 2:2 + len(toks)] = toks
    return torch.from_numpy(rec).to(device)


def unpack_hypotheses(rec: torch.Tensor, nbest: int):
    import numpy as np

    r = rec.cpu().numpy()
    out = []
    for i in range(r.shape[0]):
        if int(r[i, 0, 0]) == PAD_ROW:
            continue
        hyps = []
        for j in range(nbest):
            n = int(r[i, j, 0])
            if n < 0:
                continue
            hyps.append((r[i, j, 2:2 + n].tolist(), float(r[i, j, 1:2].view(np.float32)[0])))
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


def results_to_records(results, nbest: int, max_tokens: int, device="cpu") -> torch.Tensor:
    """Speech2Text result lists [(text, tokens, token_int, Hypothesis), ...] per utterance -> hypothesis records (token ids with sos / eos
    stripped, as the result tuples carry them, and the total score)."""
    return pack_hypotheses([[(r[2], float(r[3].score)) for r in res[:nbest]] for res in results], nbest, max_tokens, device=device)


def all_gather_results(results, nbest: int, max_tokens: int, world: int, device="cpu"):
    """The single exchange of the utterance-sharded path: every rank contributes the records of its own utterances and gets
    [rank][utterance][(token ids, score), ...] of the whole job (NCCL on GPUs, gloo in the CPU tests)."""
    rec = results_to_records(results, nbest, max_tokens, device=device)
    allrec = gather_hypotheses(rec, world)
    allrec = allrec.cpu()
    return [unpack_hypotheses(allrec[r], nbest) for r in range(world)]
