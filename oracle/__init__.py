"""CPU oracle for the ESPnet2 Speech2Text hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A plain PyTorch-CPU fp32 restatement of the reference's algorithm for the path named in
BASELINE.json (DefaultFrontend -> UtteranceMVN -> ConformerEncoder -> CTC head ->
BatchBeamSearch with TransformerDecoder + CTCPrefixScorer), written from the reference's
behaviour with every function citing the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl reference
legs may import this package, and only as the checker / CPU baseline.  Nothing under
``espnet_b200/`` imports it; the product path raises if the CUDA library is missing.

Parity pinning: the reference has no golden vectors for this path.  The oracle is
pinned against outputs of the reference itself: ``tests/golden/make_golden.py`` imports the real
``espnet2.bin.asr_inference.Speech2Text`` from /root/reference, runs it on seeded inputs and
commits the stage-boundary tensors as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
checks this package against them (and, when /root/reference is mounted,
``tests/test_oracle_vs_reference.py`` re-runs the live reference).
"""
from .frontend import slaney_mel_matrix, stft_power, log_mel, utterance_mvn, frontend_forward  # noqa: F401
from .encoder import conformer_encode, ctc_logits, ctc_greedy  # noqa: F401
from .search import OracleDecoder, CTCPrefixScorerTH, batch_beam_search  # noqa: F401
from .pipeline import OracleSpeech2Text  # noqa: F401
