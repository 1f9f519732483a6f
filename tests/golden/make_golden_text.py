"""Generate tests/golden/text.json from the UNMODIFIED reference text classes (espnet2/text/*): TokenIDConverter.ids2tokens /
tokens2ids and tokens2text of the char / word tokenizers on fixed inputs.  Run in the build container only.

    python tests/golden/make_golden_text.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()
from espnet2.text.build_tokenizer import build_tokenizer  # noqa: E402
from espnet2.text.token_id_converter import TokenIDConverter  # noqa: E402

token_list = ["<blank>", "<unk>", "<space>", "a", "b", "c", "'", "▁he", "llo", "▁wor", "ld", "<sos/eos>"]
ids = [3, 4, 2, 5, 6, 3, 2, 2, 4]
conv = TokenIDConverter(token_list)
toks = conv.ids2tokens(ids)
out = dict(token_list=token_list, ids=ids, tokens=toks, back=conv.tokens2ids(toks + ["zzz"]), nvocab=conv.get_num_vocabulary_size(),
           char_text=build_tokenizer("char").tokens2text(toks), char_repr=repr(build_tokenizer("char")),
           char_text_custom_space=build_tokenizer("char", space_symbol="a").tokens2text(toks),
           word_text=build_tokenizer("word").tokens2text(["hello", "world", "<unk>"]),
           word_text_delim=build_tokenizer("word", delimiter="|").tokens2text(["hello", "world"]), word_repr=repr(build_tokenizer("word")))
json.dump(out, open(os.path.join(HERE, "text.json"), "w"), ensure_ascii=False, indent=1)
print(out)
