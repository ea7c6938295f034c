"""CPU: cris.pytorch_b200.tokenizer (own implementation of CLIP's byte-level BPE + the reference's `tokenize` framing,
utils/dataset.py:43-84) against (a) id vectors the reference's own tokenizer produced (tests/golden/tokenizer_r02.json,
oracle/make_tokenizer_golden.py) and (b) the reference tokenizer itself, live, on generated sentences when the
reference checkout is present.  Integer ids: the bar is exact equality.  The merge table is reference DATA that is not
committed here: the tests skip when neither CRIS_BPE_VOCAB nor the staged baseline/_ref copy exists."""
import json
import os
import random
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def tok():
    from cris.pytorch_b200 import tokenizer as T
    try:
        path = T.default_vocab_path()
    except FileNotFoundError as e:
        pytest.skip(str(e))
    return T.BPETokenizer(path)


def test_vocabulary_layout(tok):
    assert len(tok.words) == 49408 and tok.sot == 49406 and tok.eot == 49407
    assert tok.ids["!"] == 0 and tok.ids["!</w>"] == 256 and len(set(tok.sym)) == 256
    assert tok.decode(tok.encode("Hello,  World")) .strip() == "hello , world"


def test_matches_reference_id_vectors(tok):
    from cris.pytorch_b200.tokenizer import tokenize
    g = json.load(open(os.path.join(HERE, "golden", "tokenizer_r02.json")))
    got17 = tokenize(g["sentences"], 17, True, tokenizer=tok)
    got77 = tokenize(g["sentences"], 77, False, tokenizer=tok)
    assert got17.dtype == torch.long and tuple(got17.shape) == (len(g["sentences"]), 17)
    assert got17.tolist() == g["len17_truncate"]
    assert got77.tolist() == g["len77"]
    long_sentence = next(s for s in g["sentences"] if s.startswith("the batter"))
    row = tokenize(long_sentence, 17, True, tokenizer=tok)[0]
    assert int(row[0]) == tok.sot and int(row[-1]) == tok.eot            # cut, last kept id replaced by EOT
    with pytest.raises(RuntimeError):
        tokenize(long_sentence, 17, False, tokenizer=tok)


def test_matches_reference_tokenizer_live(tok):
    ref_root = "/root/reference"
    if not os.path.exists(os.path.join(ref_root, "utils", "simple_tokenizer.py")):
        pytest.skip("reference checkout not present")
    sys.path[:0] = [os.path.join(HERE, "stubs"), ref_root]
    try:
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
        from utils.simple_tokenizer import SimpleTokenizer
        ref = SimpleTokenizer(os.path.join(ref_root, "utils", "bpe_simple_vocab_16e6.txt.gz"))
    finally:
        del sys.path[:2]
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
            del sys.modules[k]
    rng = random.Random(3)
    words = ("the a man woman dog's cat can't left right-most 2nd 3 #4 red blue umbrella, pizza; (blurry) w/ o'clock "
             "Zebra GIRAFFE skateboarder's snowboarding refrigerator toothbrush café naïve 100% a&b x-ray it's they'll").split()
    for _ in range(400):
        s = " ".join(rng.choice(words) for _ in range(rng.randint(1, 14)))
        assert tok.encode(s) == ref.encode(s), s
