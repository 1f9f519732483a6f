"""Command-line streaming decoding: ``python -m espnet_b200.bin_asr_inference_streaming --output_dir ... --sim_chunk_length 640
--data_path_and_name_and_type wav.scp,speech,sound --asr_train_config ... --asr_model_file ...``.

The contract of espnet2/bin/asr_inference_streaming.py:360-487 (``inference()``): every utterance of the scp is pushed through
``Speech2TextStreaming`` in chunks of ``--sim_chunk_length`` samples (``is_final=False``) followed by the remainder with ``is_final=True``
(``--sim_chunk_length 0``: the whole utterance in one final push), and the n-best of the final push is written as
``<output_dir>/{n}best_recog/{token,token_int,score,text}`` (the same writer and placeholder hypothesis for a TooShortUttError as the offline tool,
bin_asr_inference.py).  Option names are the reference's (:490-642) for everything on this path.
"""
import argparse
import logging
from typing import Dict, Optional, Sequence, Tuple, Union

import torch

from .bin_asr_inference import ResultDirWriter, iter_scp, read_sound, write_results
from .errors import TooShortUttError
from .search import Hypothesis


def simulate_stream(speech2text, speech, sim_chunk_length: int):
    """Push one utterance chunk by chunk; returns what the final push returned (asr_inference_streaming.py:462-476)."""
    speech = torch.as_tensor(speech)
    if sim_chunk_length == 0:
        return speech2text(speech, is_final=True)
    n_full = len(speech) // sim_chunk_length
    for i in range(n_full):
        speech2text(speech[i * sim_chunk_length: (i + 1) * sim_chunk_length], is_final=False)
    return speech2text(speech[n_full * sim_chunk_length:], is_final=True)


def inference(output_dir: str, data_path_and_name_and_type: Sequence[Tuple[str, str, str]], key_file: Optional[str] = None, batch_size: int = 1,
              nbest: int = 1, ngpu: int = 1, sim_chunk_length: int = 0, log_level: Union[int, str] = "INFO", speech2text=None,
              **speech2text_kwargs) -> Dict[str, list]:
    from .asr_inference_streaming import Speech2TextStreaming

    logging.basicConfig(level=log_level, format="%(asctime)s (%(module)s:%(lineno)d) %(levelname)s: %(message)s")
    if batch_size > 1:
        raise NotImplementedError("batch decoding is not implemented")           # as the reference (:393-394)
    if ngpu < 1 and speech2text is None:
        raise RuntimeError("espnet_b200 has no CPU path: --ngpu must be >= 1")
    scps = [(p, name, typ) for p, name, typ in data_path_and_name_and_type if name == "speech"]
    if len(scps) != 1 or scps[0][2] != "sound":
        raise NotImplementedError("exactly one --data_path_and_name_and_type <scp>,speech,sound is expected")
    if speech2text is None:
        speech2text = Speech2TextStreaming(nbest=nbest, device="cuda", **speech2text_kwargs)
    placeholder = [(" ", ["<space>"], [2], Hypothesis(score=0.0, scores={}, states={}, yseq=torch.zeros(0, dtype=torch.long)))] * nbest
    out: Dict[str, list] = {}
    with ResultDirWriter(output_dir) as writer:
        for key, path in iter_scp(scps[0][0], key_file):
            try:
                results = simulate_stream(speech2text, read_sound(path), sim_chunk_length)
            except TooShortUttError as e:
                logging.warning(f"Utterance {key} {e}")
                speech2text.reset()
                results = placeholder
            write_results(writer, key, results, nbest)
            out[key] = results
    return out


def get_parser():
    s2b = lambda v: str(v).lower() in ("1", "true", "yes", "y")  # noqa: E731
    none_or = lambda v: None if str(v).lower() in ("none", "null", "nil", "") else v  # noqa: E731
    p = argparse.ArgumentParser(description="Streaming ASR Decoding (espnet_b200)", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("--log_level", type=lambda x: x.upper(), default="INFO")
    p.add_argument("--output_dir", type=str, required=True)
    p.add_argument("--ngpu", type=int, default=1)
    p.add_argument("--dtype", default="float32", choices=["float32"])
    p.add_argument("--data_path_and_name_and_type", type=lambda v: tuple(v.split(",")), required=True, action="append")
    p.add_argument("--key_file", type=none_or)
    p.add_argument("--sim_chunk_length", type=int, default=0, help="The length of one chunk, to which speech will be divided for evalution of streaming processing.")
    p.add_argument("--asr_train_config", type=str, required=True)
    p.add_argument("--asr_model_file", type=str, required=True)
    p.add_argument("--lm_train_config", type=none_or)
    p.add_argument("--lm_file", type=none_or)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--nbest", type=int, default=1)
    p.add_argument("--beam_size", type=int, default=20)
    p.add_argument("--penalty", type=float, default=0.0)
    p.add_argument("--maxlenratio", type=float, default=0.0)
    p.add_argument("--minlenratio", type=float, default=0.0)
    p.add_argument("--ctc_weight", type=float, default=0.5)
    p.add_argument("--lm_weight", type=float, default=1.0)
    p.add_argument("--normalize_length", type=s2b, default=False)
    p.add_argument("--disable_repetition_detection", type=s2b, default=False)
    p.add_argument("--encoded_feat_length_limit", type=int, default=0)
    p.add_argument("--decoder_text_length_limit", type=int, default=0)
    p.add_argument("--token_type", type=none_or, default=None, choices=["char", "bpe", "word", None])
    p.add_argument("--bpemodel", type=none_or, default=None)
    return p


def main(cmd=None):
    inference(**vars(get_parser().parse_args(cmd)))


if __name__ == "__main__":
    main()
