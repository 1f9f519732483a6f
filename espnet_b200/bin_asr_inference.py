"""Command-line decoding: ``python -m espnet_b200.bin_asr_inference --output_dir ... --data_path_and_name_and_type wav.scp,speech,sound ...``.

The output-side contract of espnet2/bin/asr_inference.py:711-906 (``inference()``) and espnet2/fileio/datadir_writer.py: a Kaldi-style result
directory ``<output_dir>/{n}best_recog/{token,token_int,score,text}`` with one ``<utterance id> <value>`` line per utterance, a TooShortUttError
turned into the placeholder hypothesis (asr_inference.py:850-856), the option names of the reference's parser (:909-1168) for everything on this
path, and ``utils/calculate_rtf.py``-compatible log lines (Speech2Text logs "speech length" / "best hypo").  What differs by design: utterances
are decoded ``--batch_size`` at a time in ONE device pass (the reference refuses batch_size > 1, :761), and with ``torchrun`` every rank decodes
the keys ``rank::world_size`` and writes ``<output_dir>/rank<r>/`` (the reference splits the key file across processes, asr.sh:1591-1618).
"""
import argparse
import logging
import os
import struct
import sys
import wave
from pathlib import Path
from typing import Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .errors import TooShortUttError
from .search import Hypothesis


class ResultDirWriter:
    """Directory tree of ``key value`` text files (the behaviour of espnet2.fileio.datadir_writer.DatadirWriter): ``w["1best_recog"]["text"]["utt1"]
    = "hello"`` appends ``utt1 hello`` to <root>/1best_recog/text.  A node is either a directory (indexed by name) or a file (assigned by key)."""

    def __init__(self, root: Union[Path, str]):
        self.path, self._dirs, self._fd, self.keys = Path(root), {}, None, set()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __getitem__(self, name: str) -> "ResultDirWriter":
        if self._fd is not None:
            raise RuntimeError("This writer points out a file")
        if name not in self._dirs:
            self._dirs[name] = ResultDirWriter(self.path / name)
        return self._dirs[name]

    def __setitem__(self, key: str, value: str):
        if self._dirs:
            raise RuntimeError("This writer points out a directory")
        if key in self.keys:
            logging.warning(f"Duplicated: {key}")
        if self._fd is None:
            self.path.parent.mkdir(parents=True, exist_ok=True)
            self._fd = self.path.open("w", encoding="utf-8")
        self.keys.add(key)
        self._fd.write(f"{key} {value}\n")

    def close(self):
        prev = None
        for child in self._dirs.values():
            child.close()
            if prev is not None and not child._dirs and not prev._dirs and prev.keys != child.keys:
                logging.warning(f"Ids are mismatching between {prev.path} and {child.path}")
            prev = child
        if self._fd is not None:
            self._fd.close()
            self._fd = None


def read_sound(path: str) -> np.ndarray:
    """A waveform as float32 in [-1, 1): RIFF/WAVE PCM 16 / 32 bit or IEEE float (first channel), or a ``.npy`` array."""
    if path.endswith(".npy"):
        return np.load(path).astype(np.float32).reshape(-1)
    with wave.open(path, "rb") as f:
        nch, width, n = f.getnchannels(), f.getsampwidth(), f.getnframes()
        raw = f.readframes(n)
    if width == 2:
        x = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        x = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    else:
        raise NotImplementedError(f"{path}: {8 * width}-bit PCM")
    return x.reshape(-1, nch)[:, 0].copy() if nch > 1 else x


def iter_scp(scp: str, key_file: Optional[str] = None, rank: int = 0, world: int = 1) -> Iterator[Tuple[str, str]]:
    """(utterance id, path) of a Kaldi ``wav.scp`` (``id path`` per line), restricted to ``key_file`` ids if given, sharded ``rank::world``."""
    keep = None
    if key_file is not None:
        with open(key_file, encoding="utf-8") as f:
            keep = {ln.split(maxsplit=1)[0] for ln in f if ln.strip()}
    i = 0
    with open(scp, encoding="utf-8") as f:
        for ln in f:
            if not ln.strip():
                continue
            key, path = ln.rstrip("\n").split(maxsplit=1)
            if keep is not None and key not in keep:
                continue
            if i % world == rank:
                yield key, path.strip()
            i += 1


def write_results(writer: ResultDirWriter, key: str, results, nbest: int):
    """asr_inference.py:884-896: token / token_int / score (/ text) of the n-best list."""
    for n, (text, token, token_int, hyp) in zip(range(1, nbest + 1), results):
        w = writer[f"{n}best_recog"]
        w["token"][key] = " ".join(token)
        w["token_int"][key] = " ".join(map(str, token_int))
        w["score"][key] = str(float(hyp.score))
        if text is not None:
            w["text"][key] = text


def inference(output_dir: str, data_path_and_name_and_type: Sequence[Tuple[str, str, str]], key_file: Optional[str] = None, batch_size: int = 1,
              nbest: int = 1, ngpu: int = 1, log_level: Union[int, str] = "INFO", speech2text=None, **speech2text_kwargs) -> Dict[str, list]:
    """Decode every utterance of the ``sound`` scp and write the result directory.  Returns {utterance id: n-best results}."""
    from .asr_inference import Speech2Text

    logging.basicConfig(level=log_level, format="%(asctime)s (%(module)s:%(lineno)d) %(levelname)s: %(message)s")
    if ngpu < 1:
        raise RuntimeError("espnet_b200 has no CPU path: --ngpu must be >= 1")
    scps = [(p, name, typ) for p, name, typ in data_path_and_name_and_type if name == "speech"]
    if len(scps) != 1 or scps[0][2] != "sound":
        raise NotImplementedError("exactly one --data_path_and_name_and_type <scp>,speech,sound is expected")
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        output_dir = os.path.join(output_dir, f"rank{rank}")
    if speech2text is None:
        speech2text = Speech2Text.from_pretrained(nbest=nbest, device="cuda", **speech2text_kwargs)
    out: Dict[str, list] = {}
    placeholder = [(" ", ["<space>"], [2], Hypothesis(score=0.0, scores={}, states={}, yseq=torch.zeros(0, dtype=torch.long)))] * nbest

    def flush(keys: List[str], waves: List[np.ndarray], writer):
        if not keys:
            return
        try:
            results = speech2text.batch_decode(waves)
        except TooShortUttError:          # one short utterance must not take the batch down: decode one by one (asr_inference.py:850-856)
            results = []
            for k, wv in zip(keys, waves):
                try:
                    results.append(speech2text(wv))
                except TooShortUttError as e:
                    logging.warning(f"Utterance {k} {e}")
                    results.append(placeholder)
        for k, res in zip(keys, results):
            write_results(writer, k, res, nbest)
            out[k] = res

    with ResultDirWriter(output_dir) as writer:
        keys, waves = [], []
        for key, path in iter_scp(scps[0][0], key_file, rank, world):
            keys.append(key); waves.append(read_sound(path))
            if len(keys) == batch_size:
                flush(keys, waves, writer)
                keys, waves = [], []
        flush(keys, waves, writer)
    return out


def get_parser():
    """The options of espnet2/bin/asr_inference.py:909-1168 that exist on this path (same names, same defaults)."""
    s2b = lambda v: str(v).lower() in ("1", "true", "yes", "y")  # noqa: E731
    none_or = lambda v: None if str(v).lower() in ("none", "null", "nil", "") else v  # noqa: E731
    p = argparse.ArgumentParser(description="ASR Decoding (espnet_b200)", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--log_level", type=lambda x: x.upper(), default="INFO")
    p.add_argument("--output_dir", type=str, required=True)
    p.add_argument("--ngpu", type=int, default=1)
    p.add_argument("--dtype", default="float32", choices=["float32"])
    p.add_argument("--data_path_and_name_and_type", type=lambda v: tuple(v.split(",")), required=True, action="append")
    p.add_argument("--key_file", type=none_or)
    p.add_argument("--asr_train_config", type=str)
    p.add_argument("--asr_model_file", type=str)
    p.add_argument("--lm_train_config", type=none_or)
    p.add_argument("--lm_file", type=none_or)
    p.add_argument("--model_tag", type=none_or)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--nbest", type=int, default=1)
    p.add_argument("--beam_size", type=int, default=20)
    p.add_argument("--penalty", type=float, default=0.0)
    p.add_argument("--maxlenratio", type=float, default=0.0)
    p.add_argument("--minlenratio", type=float, default=0.0)
    p.add_argument("--ctc_weight", type=float, default=0.5)
    p.add_argument("--lm_weight", type=float, default=1.0)
    p.add_argument("--normalize_length", type=s2b, default=False)
    p.add_argument("--token_type", type=none_or, default=None, choices=["char", "bpe", "word", None])
    p.add_argument("--bpemodel", type=none_or, default=None)
    return p


def main(cmd=None):
    args = vars(get_parser().parse_args(cmd))
    inference(**args)


if __name__ == "__main__":
    main()
