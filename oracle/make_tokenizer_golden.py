"""Writes tests/golden/tokenizer_r02.json: token ids produced by the REFERENCE's own `tokenize` (utils/dataset.py:43-84 over
utils/simple_tokenizer.py) for a fixed list of sentences.  Run in the build container (needs /root/reference; `lmdb` and
`ftfy`, which are not installed here, are the import-time stubs of tests/stubs — ftfy.fix_text is the identity on these
sentences' ASCII / well-formed UTF-8):  python -m oracle.make_tokenizer_golden"""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(REPO, "tests", "golden", "tokenizer_r02.json")

SENTENCES = [
    "the man in the red shirt", "A woman's   dog, running!", "2 zebras &amp; 13 giraffes", "left-most person (blurry)",
    "naïve café — déjà vu", "umbrella's handle isn't visible", "  THE BIG   WHITE   PLANE  ", "girl in pink",
    "person holding a hot-dog w/ mustard #3", "front row 2nd from right", "guy", "", "!!!", "it's they're we've i'm he'll she'd",
    "catcher", "the batter in the white uniform swinging at the ball near home plate while the umpire watches closely behind",
    "giraffe on the left, taller one", "3rd donut from top-left; chocolate w/ sprinkles", "blue&white striped umbrella",
    "bowl of broccoli 12 o'clock", "laptop @ right edge", "woman's hand w/ ring", "pizza slice closest 2 us",
    "the 100% wool sweater", "man\twith\ttabs\nand newlines", "ÉCOLE élève Ünïcödé straße", "日本語 のテキスト", "emoji 😀 face",
    "a" * 40, "skier in yellow jacket, far right", "bottom left corner sandwich half", "kid wearing #7 jersey",
    "&lt;tag&gt; &amp;amp; entities", "white car behind the bus", "second elephant from the left", "top shelf, 3 books",
    "zebra w/ head down", "lady in black dress holding wine glass", "partial person at very edge of pic on right",
    "the clock tower's face", "don't pick the dog; pick the cat", "chair - empty one", "bear.", "bear?", "bear...",
]


def main():
    sys.path.insert(0, os.path.join(REPO, "tests", "stubs"))
    sys.path.insert(0, "/root/reference")
    from utils.dataset import tokenize
    out = {"sentences": SENTENCES,
           "len17_truncate": [tokenize(s, 17, True).squeeze(0).tolist() for s in SENTENCES],
           "len77": [tokenize(s, 77, False).squeeze(0).tolist() for s in SENTENCES]}
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
