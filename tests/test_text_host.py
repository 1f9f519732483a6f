"""Output-side text conversion (host only) against the reference's own classes (tests/golden/text.json, made by
tests/golden/make_golden_text.py from espnet2/text/*) and against sentencepiece itself for the bpe tokenizer."""
import argparse
import json
import os

import pytest

from espnet_b200.text import TokenIDConverter, build_tokenizer, tokenizer_for_inference

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "text.json"), encoding="utf-8"))


def test_token_id_converter_matches_reference():
    conv = TokenIDConverter(G["token_list"])
    assert conv.ids2tokens(G["ids"]) == G["tokens"]
    assert conv.tokens2ids(G["tokens"] + ["zzz"]) == G["back"]        # unknown token -> <unk>
    assert conv.get_num_vocabulary_size() == G["nvocab"]
    with pytest.raises(RuntimeError):
        TokenIDConverter(["a", "a", "<unk>"])                          # duplicated symbol (token_id_converter.py:36-38)
    with pytest.raises(RuntimeError):
        TokenIDConverter(["a", "b"])                                   # no <unk> (token_id_converter.py:41-45)


def test_token_list_file(tmp_path):
    f = tmp_path / "tokens.txt"
    f.write_text("\n".join(G["token_list"]) + "\n", encoding="utf-8")
    assert TokenIDConverter(str(f)).token_list == G["token_list"]


def test_char_and_word_tokens2text_match_reference():
    assert build_tokenizer("char").tokens2text(G["tokens"]) == G["char_text"]
    assert build_tokenizer("char", space_symbol="a").tokens2text(G["tokens"]) == G["char_text_custom_space"]
    assert build_tokenizer("word").tokens2text(["hello", "world", "<unk>"]) == G["word_text"]
    assert build_tokenizer("word", delimiter="|").tokens2text(["hello", "world"]) == G["word_text_delim"]
    assert repr(build_tokenizer("word")) == G["word_repr"]


def test_bpe_tokens2text_is_sentencepiece_decode(tmp_path):
    spm = pytest.importorskip("sentencepiece")
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(["hello world", "the quick brown fox", "jumps over the lazy dog", "hello there world"] * 20), encoding="utf-8")
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(tmp_path / "bpe"), vocab_size=40, model_type="bpe",
                                   hard_vocab_limit=False, minloglevel=2)
    model = str(tmp_path / "bpe.model")
    sp = spm.SentencePieceProcessor()
    sp.load(model)
    pieces = sp.EncodeAsPieces("hello brown dog")
    tok = build_tokenizer("bpe", model)
    assert tok.tokens2text(pieces) == sp.DecodePieces(pieces) == "hello brown dog"
    with pytest.raises(ValueError):
        build_tokenizer("bpe", None)


def test_inference_selection_rule():
    """asr_inference.py:395-430: arguments override the training config; bpe without a model and token_type None give no tokenizer."""
    args = argparse.Namespace(token_type="char", bpemodel=None)
    assert type(tokenizer_for_inference(None, None, args)).__name__ == "CharTokenizer"
    assert type(tokenizer_for_inference("word", None, args)).__name__ == "WordTokenizer"
    assert tokenizer_for_inference(None, None, argparse.Namespace(token_type=None, bpemodel=None)) is None
    assert tokenizer_for_inference(None, None, argparse.Namespace(token_type="bpe", bpemodel=None)) is None
    assert tokenizer_for_inference(None, None, argparse.Namespace()) is None
    with pytest.raises(NotImplementedError):
        tokenizer_for_inference("phn", None, args)


def test_end_of_search_log_lines(caplog):
    """The lines utils/calculate_rtf.py keys on ('best hypo') and the score summary of beam_search.py:476-487."""
    import logging
    import types

    import torch

    import espnet_b200
    from espnet_b200.search import Hypothesis

    s2t = espnet_b200.Speech2Text.__new__(espnet_b200.Speech2Text)
    s2t.converter = TokenIDConverter(G["token_list"])
    s2t.beam_search = types.SimpleNamespace(weights=dict(decoder=0.7, ctc=0.3, length_bonus=0.0))
    s2t.tokenizer, s2t.nbest = build_tokenizer("char"), 1
    hyp = Hypothesis(yseq=torch.tensor([11, 3, 4, 2, 5, 11]), score=-3.25, scores=dict(decoder=-2.0, ctc=-6.0))
    with caplog.at_level(logging.INFO, logger="espnet_b200.asr_inference"):
        res = s2t._results([hyp])
    assert res[0][0] == "ab c" and res[0][1] == ["a", "b", "<space>", "c"] and res[0][2] == [3, 4, 2, 5]
    text = caplog.text
    assert "total log probability: -3.25" in text and "normalized log probability: -0.54" in text
    assert "total number of ended hypotheses: 1" in text and "best hypo: ab<space>c" in text
    assert " -2.00 * 0.7 =  -1.40 for decoder" in text


def test_speech2text_refuses_unimplemented_reference_keywords():
    """ngram_file / transducer_conf / streaming ... change the decoding algorithm: they must raise, not be swallowed (and they are
    checked before any device work, so this runs without a GPU)."""
    import pytest

    import espnet_b200

    for kw in (dict(ngram_file="4gram.bin"), dict(transducer_conf={"search_type": "default"}), dict(streaming=True),
               dict(quantize_asr_model=True)):
        with pytest.raises(NotImplementedError, match=list(kw)[0]):
            espnet_b200.Speech2Text(asr_model=object(), **kw)
    # defaults of those keywords pass this check (the next failure is the missing CUDA device / library, not the keyword filter)
    with pytest.raises(Exception) as e:
        espnet_b200.Speech2Text(asr_model=None, asr_train_config=None, lm_file=None, streaming=False, quantize_modules=["Linear"], device="cpu")
    assert "does not implement" not in str(e.value)


def test_result_dir_writer_and_scp_reader(tmp_path):
    """The result directory of the CLI (asr_inference.py:884-896 through DatadirWriter semantics) and the wav.scp / key_file / rank sharding."""
    import wave

    import numpy as np
    import torch

    from espnet_b200.bin_asr_inference import ResultDirWriter, iter_scp, read_sound, write_results
    from espnet_b200.search import Hypothesis

    with ResultDirWriter(tmp_path / "out") as w:
        for key, toks in (("utt1", [3, 4]), ("utt2", [5])):
            res = [("t e", [f"t{t}" for t in toks], toks, Hypothesis(yseq=torch.tensor([9] + toks + [9]), score=-1.5))]
            write_results(w, key, res, nbest=1)
        with pytest.raises(RuntimeError):
            w["1best_recog"]["x"] = "a directory cannot be assigned"
    assert (tmp_path / "out/1best_recog/token").read_text() == "utt1 t3 t4\nutt2 t5\n"
    assert (tmp_path / "out/1best_recog/token_int").read_text() == "utt1 3 4\nutt2 5\n"
    assert (tmp_path / "out/1best_recog/score").read_text() == "utt1 -1.5\nutt2 -1.5\n"
    assert (tmp_path / "out/1best_recog/text").read_text() == "utt1 t e\nutt2 t e\n"
    x = (np.sin(np.arange(1600) / 10.0) * 20000).astype("<i2")
    with wave.open(str(tmp_path / "a.wav"), "wb") as f:
        f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(x.tobytes())
    np.save(tmp_path / "b.npy", np.ones(10, dtype=np.float32))
    (tmp_path / "wav.scp").write_text(f"u1 {tmp_path / 'a.wav'}\nu2 {tmp_path / 'b.npy'}\nu3 {tmp_path / 'a.wav'}\n")
    (tmp_path / "keys").write_text("u1\nu3 something\n")
    assert [k for k, _ in iter_scp(str(tmp_path / "wav.scp"))] == ["u1", "u2", "u3"]
    assert [k for k, _ in iter_scp(str(tmp_path / "wav.scp"), str(tmp_path / "keys"))] == ["u1", "u3"]
    assert [k for k, _ in iter_scp(str(tmp_path / "wav.scp"), None, rank=1, world=2)] == ["u2"]
    np.testing.assert_allclose(read_sound(str(tmp_path / "a.wav")), x.astype(np.float32) / 32768.0)
    assert read_sound(str(tmp_path / "b.npy")).shape == (10,)


def test_cli_inference_batching_writer_and_placeholder_with_a_stub_engine(tmp_path):
    """bin_asr_inference.inference host logic (asr_inference.py:824-906) without a device: scp / key-file reading, --batch_size grouping, the
    per-utterance retry when a batch holds a too-short utterance and the reference's placeholder hypothesis, the {n}best_recog/{token,token_int,score,text}
    files.  The engine is a stub with Speech2Text's batch_decode / __call__ contract."""
    import wave as wavmod

    import numpy as np
    import torch

    from espnet_b200 import Hypothesis, TooShortUttError
    from espnet_b200.bin_asr_inference import inference, iter_scp, read_sound

    class Engine:
        calls = []

        def _one(self, w):
            if len(w) < 1000:
                raise TooShortUttError("too short", len(w), 1000)
            n = len(w) // 1000
            hyp = Hypothesis(yseq=torch.tensor([9] + list(range(3, 3 + n)) + [9]), score=torch.tensor(-float(n)))
            return [(f"text{n}", [f"t{i}" for i in range(n)], list(range(3, 3 + n)), hyp), (None, ["x"], [4], hyp)]

        def batch_decode(self, waves):
            self.calls.append(len(waves))
            return [self._one(w) for w in waves]

        def __call__(self, w):
            return self._one(w)

    lens = {"a": 3000, "b": 500, "c": 5000, "d": 2000, "e": 4000}
    lines = []
    for k, n in lens.items():
        pcm = (np.arange(n) % 100).astype("<i2")
        with wavmod.open(str(tmp_path / f"{k}.wav"), "wb") as f:
            f.setnchannels(1); f.setsampwidth(2); f.setframerate(16000); f.writeframes(pcm.tobytes())
        lines.append(f"{k} {tmp_path / (k + '.wav')}")
    (tmp_path / "wav.scp").write_text("\n".join(lines) + "\n")
    (tmp_path / "keys").write_text("a\nb\nc\ne\n")
    assert [k for k, _ in iter_scp(str(tmp_path / "wav.scp"), str(tmp_path / "keys"))] == ["a", "b", "c", "e"]
    assert [k for k, _ in iter_scp(str(tmp_path / "wav.scp"), None, rank=1, world=2)] == ["b", "d"]
    x = read_sound(str(tmp_path / "a.wav"))
    assert x.dtype == np.float32 and len(x) == 3000 and abs(float(x[5]) - 5 / 32768.0) < 1e-9
    eng = Engine()
    out = inference(str(tmp_path / "dec"), [(str(tmp_path / "wav.scp"), "speech", "sound")], key_file=str(tmp_path / "keys"), batch_size=3, nbest=2,
                    speech2text=eng)
    assert eng.calls == [3, 1] and list(out) == ["a", "b", "c", "e"]            # one batch of three (holding the short one), then the rest
    read = lambda p: dict((ln.split(maxsplit=1) + [""])[:2] for ln in (tmp_path / "dec" / p).read_text().splitlines())  # noqa: E731
    tok, txt, sc = read("1best_recog/token_int"), read("1best_recog/text"), read("1best_recog/score")
    assert tok["a"].split() == ["3", "4", "5"] and tok["b"].strip() == "2" and txt["c"].strip() == "text5" and float(sc["e"]) == -4.0
    assert read("1best_recog/token")["b"].strip() == "<space>" and float(sc["b"]) == 0.0        # the placeholder of asr_inference.py:850-856
    assert list(read("2best_recog/text")) == ["b"] and read("2best_recog/token")["a"].strip() == "x"    # text is None for the stub's 2nd best; the placeholder has one
    import pytest

    with pytest.raises(RuntimeError):
        inference(str(tmp_path / "dec2"), [(str(tmp_path / "wav.scp"), "speech", "sound")], ngpu=0, speech2text=eng)
