"""-m gpu parity at the BENCHMARKED configuration (BASELINE.json configs[1] / bench.py `conformer_large_joint_64x30s`):
Conformer-large 12L/512d/8h (ff 2048, kernel 31) + 6L Transformer decoder, V=5000, 30-s utterances (T = 937 encoder frames) plus one
ragged 15-s utterance in the same batch, joint CTC/attention decoding with beam 10 -- against the CPU oracle (pinned to the reference by
tests/test_oracle_golden.py).  These are the shapes the throughput number is measured on: 937-key attention rows, the 256-column GEMM
variants at M = B*937, the conv2 implicit GEMM at T1 = 1875, cross-attention over 937 memory frames.

Tolerances: encoder output atol 1e-4 per utterance (O(1) activations after the final LayerNorm); CTC-greedy ids identical wherever the
oracle's own top-2 logit margin exceeds 20x the measured logit error (random-init models have near-ties); n-best token sequences
identical and scores within rtol 2e-4 (the reference's Batch- vs non-batch search tolerance is rtol 1e-6 on identical arithmetic,
test/espnet2/legacy/test_batch_beam_search.py:190-195; its cached-vs-uncached decoder tolerance is 1e-4, test_transformer_decode.py:9)."""
import pytest
import torch

import oracle
from oracle import encoder as OE
from gpu_util import random_weights, refbuild, speech2text

pytestmark = pytest.mark.gpu

LARGE = dict(d_model=512, heads=8, ff=2048, enc_layers=12, dec_layers=6, vocab=5000, kernel=31)


@pytest.fixture(scope="module")
def large():
    torch.set_num_threads(min(16, torch.get_num_threads()))
    w = random_weights(LARGE, seed=0)
    lens = [480000, 480000, 240000]
    waves = [refbuild.waveform(300 + i, n) for i, n in enumerate(lens)]
    return w, waves


def _maxerr(a, b):
    return (a.double().cpu() - torch.as_tensor(b).double()).abs().max().item()


def test_large_encoder_and_ctc_greedy_vs_oracle(large):
    w, waves = large
    s2t = speech2text(LARGE, w, beam_size=10, ctc_weight=0.3)
    o = oracle.OracleSpeech2Text(LARGE, w, beam_size=10, ctc_weight=0.3)
    speech, sl = s2t._to_batch(waves)
    enc, elens = s2t.asr_model.encode(speech, sl)
    lg = s2t.asr_model.ctc.logits(enc, s2t.asr_model.enc_split(enc))
    greedy = s2t.ctc_greedy(waves)
    assert elens.tolist() == [937, 937, 468]
    for i, wv in enumerate(waves):
        ref = o.encode(wv)
        assert ref.shape[0] == int(elens[i])
        e = _maxerr(enc[i, : ref.shape[0]], ref)
        ref_lg = OE.ctc_logits(ref, o.w)
        el = _maxerr(lg[i, : ref.shape[0]], ref_lg)
        top2 = ref_lg.topk(2, dim=-1)[0]
        gap = top2[:, 0] - top2[:, 1]
        print(f"utt{i} T={ref.shape[0]}: encoder max abs err {e:.3e}, logits max abs err {el:.3e}, min top-2 margin {gap.min().item():.3e}")
        assert e < 1e-4
        assert el < 2e-4
        _, ids = OE.ctc_greedy(ref, o.w)
        if gap.min().item() > 20 * el:
            assert greedy[i] == ids.tolist()
        safe = gap > 20 * el
        am_got = lg[i, : ref.shape[0]].argmax(-1).cpu()
        assert bool((ref_lg.argmax(-1)[safe] == am_got[safe]).all())
        assert int(safe.sum()) > 0.9 * ref.shape[0]


@pytest.mark.parametrize("steps", [8, 16])
def test_large_joint_beam10_vs_oracle(large, steps):
    w, waves = large
    kw = dict(beam_size=10, ctc_weight=0.3, maxlenratio=-float(steps), nbest=5)
    s2t = speech2text(LARGE, w, **kw)
    o = oracle.OracleSpeech2Text(LARGE, w, **kw)
    res = s2t.batch_decode(waves)
    for i in ((0, 2) if steps == 16 else (0, 1, 2)):     # 16 oracle steps cost ~4 s of CPU per utterance
        ref = o(waves[i])
        assert len(res[i]) == len(ref) > 0
        for a, b in zip(res[i], ref):
            print(f"steps={steps} utt{i}: {a[3].yseq.tolist()} {a[3].score:.5f} vs {b[3].score:.5f}")
            assert a[3].yseq.tolist() == b[3].yseq.tolist()
            assert abs(a[3].score - b[3].score) <= 2e-4 * max(1.0, abs(b[3].score))
