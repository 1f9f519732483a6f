"""Host post-processing of the device search (back-pointer walk, eos at maxlen, sort): vectorised version vs a plain loop."""
import numpy as np
import pytest
import torch

from espnet_b200.search import BatchBeamSearch


def _loop_collect(bs, U, W, steps, maxlen, bpp, bpt, cnt, es, el, sc, sd, sct):
    """beam_search.py:452-459 / batch_beam_search.py:392-407 restated with per-hypothesis loops."""
    out = []
    for u in range(U):
        hyps = []
        for e in range(int(cnt[u])):
            step, slot = int(es[u, e]), int(el[u, e])
            toks, s = [], slot
            for j in range(step, -1, -1):
                toks.append(int(bpt[j, s]))
                s = int(bpp[j, s])
            toks.reverse()
            yseq = [bs.sos] + toks
            if step == int(maxlen[u]) - 1:
                yseq.append(bs.eos)
            hyps.append((yseq, float(sc[u, e]), float(sd[u, e]), float(sct[u, e]), step))
        if bs.normalize_length:
            hyps.sort(key=lambda h: h[1] / (len(h[0]) - 1), reverse=True)
        else:
            hyps.sort(key=lambda h: h[1], reverse=True)
        out.append(hyps)
    return out


@pytest.mark.parametrize("normalize_length,penalty", [(False, 0.0), (True, 0.5)])
def test_collect_matches_loop(normalize_length, penalty):
    rng = np.random.RandomState(0)
    U, W, steps, V = 5, 4, 9, 30
    n = U * W
    bs = BatchBeamSearch.__new__(BatchBeamSearch)
    bs.sos = bs.eos = V - 1
    bs.decoder, bs.ctc, bs.penalty, bs.normalize_length = object(), object(), penalty, normalize_length
    bpt = rng.randint(0, V - 1, size=(steps, n)).astype(np.int32)
    bpp = np.stack([(np.arange(n) // W) * W + rng.randint(0, W, size=n) for _ in range(steps)]).astype(np.int32)
    maxlen = torch.tensor([9, 6, 9, 4, 9])
    cap_e = W * steps
    cnt = np.array([7, 0, 3, cap_e, 1], dtype=np.int32)
    es = np.zeros((U, cap_e), np.int32); el = np.zeros((U, cap_e), np.int32)
    for u in range(U):
        es[u] = rng.randint(0, int(maxlen[u]), size=cap_e)
        el[u] = u * W + rng.randint(0, W, size=cap_e)
    sc = rng.randn(U, cap_e).astype(np.float32); sc[0, 2] = sc[0, 5]      # a tie keeps the ended order
    sd, sct = rng.randn(U, cap_e).astype(np.float32), rng.randn(U, cap_e).astype(np.float32)
    t = torch.from_numpy
    got = bs._collect(U, W, steps, maxlen, t(bpp), t(bpt), t(cnt), t(es), t(el), t(sc), t(sd), t(sct))
    ref = _loop_collect(bs, U, W, steps, maxlen, bpp, bpt, cnt, es, el, sc, sd, sct)
    assert len(got) == U
    for u in range(U):
        assert len(got[u]) == len(ref[u]) == int(cnt[u])
        for h, (yseq, score, dec, ctc, step) in zip(got[u], ref[u]):
            assert h.yseq.tolist() == yseq and h.yseq.dtype == torch.int64
            assert h.score == score and h.scores["decoder"] == dec and h.scores["ctc"] == ctc
            assert ("length_bonus" in h.scores) == (penalty != 0)
            if penalty != 0:
                assert h.scores["length_bonus"] == float(step + 1)


def test_group_bounds_partition_the_batch(monkeypatch):
    """Utterance groups (BatchBeamSearch.n_groups) are contiguous, cover the batch once and are only used for large batches."""
    bs = BatchBeamSearch.__new__(BatchBeamSearch)
    monkeypatch.delenv("ESPB_SEARCH_GROUPS", raising=False)
    bs.group_min_utts, bs.n_groups = 16, 1
    assert bs._group_bounds(64) == [(0, 64)]
    bs.n_groups = 2
    assert bs._group_bounds(8) == [(0, 8)]                      # below group_min_utts
    assert bs._group_bounds(64) == [(0, 32), (32, 64)]
    bs.n_groups = 3
    b = bs._group_bounds(17)
    assert b == [(0, 6), (6, 12), (12, 17)] and all(u1 > u0 for u0, u1 in b)
    bs.group_min_utts = 2
    assert bs._group_bounds(5) == [(0, 5)]                      # fewer than 2 utterances per group -> no split
    monkeypatch.setenv("ESPB_SEARCH_GROUPS", "4")
    assert bs._group_bounds(8) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    monkeypatch.setenv("ESPB_SEARCH_GROUPS", "1")
    assert bs._group_bounds(64) == [(0, 64)]


def test_search_rejects_unsupported_setups():
    import espnet_b200

    dec = espnet_b200.TransformerDecoder(50, 64, attention_heads=4, linear_units=128, num_blocks=1)
    ctc = espnet_b200.CTC(50, 64)
    with pytest.raises(NotImplementedError):       # beam wider than the selection kernel's capacity (64)
        BatchBeamSearch(dict(decoder=dec, ctc=ctc), dict(decoder=0.7, ctc=0.3), 65, 50, 49, 49, pre_beam_score_key="full")
    with pytest.raises(NotImplementedError):       # joint decoding needs vocab > 1.5 * beam (pre-beam)
        BatchBeamSearch(dict(decoder=dec, ctc=ctc), dict(decoder=0.7, ctc=0.3), 40, 50, 49, 49, pre_beam_score_key="full")
    with pytest.raises(ValueError):                # all scorer weights zero
        BatchBeamSearch(dict(decoder=dec, ctc=ctc), dict(decoder=0.0, ctc=0.0), 4, 50, 49, 49)
    bs = BatchBeamSearch(dict(decoder=dec, ctc=ctc), dict(decoder=0.7, ctc=0.3, length_bonus=0.5), 4, 50, 49, 49, pre_beam_score_key="full")
    assert bs.pre_beam_size == 6 and bs.do_pre_beam and bs.penalty == 0.5
    assert BatchBeamSearch(dict(decoder=dec, ctc=ctc), dict(decoder=1.0, ctc=0.0), 4, 50, 49, 49).ctc is None   # zero-weight scorer dropped
