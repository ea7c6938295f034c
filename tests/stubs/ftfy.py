"""Test stub: utils/simple_tokenizer.py:6 imports ftfy; only fix_text is used (for real captions)."""


def fix_text(text):
    return text
