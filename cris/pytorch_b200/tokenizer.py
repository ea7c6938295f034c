"""Sentence -> token ids, the text half of the reference's input side (utils/dataset.py:43-84 `tokenize`, which wraps
utils/simple_tokenizer.py: CLIP's lower-cased byte-level BPE).  Host code: the ids feed `CRIS.forward(img, word, mask)`.

Own implementation of the published algorithm (byte-level BPE of CLIP / GPT-2):
  * text cleaning: (ftfy.fix_text when ftfy is installed,) html.unescape twice, strip, runs of whitespace -> one space,
    lower case;
  * split with CLIP's pattern (specials | 's 't 're 've 'm 'll 'd | letter runs | single digits | other non-space runs);
  * each piece -> its UTF-8 bytes -> one printable symbol per byte, the last symbol tagged '</w>';
  * repeatedly fuse the adjacent symbol pair with the lowest merge rank (all its occurrences) until none is ranked;
  * ids: 256 byte symbols, 256 byte symbols + '</w>', the 48 894 merges in file order, then <|startoftext|> = 49406 and
    <|endoftext|> = 49407.

The merge table is DATA of the reference checkout (utils/bpe_simple_vocab_16e6.txt.gz) and is not part of this
repository: pass its path, or set CRIS_BPE_VOCAB, or stage the reference under baseline/_ref (tools/stage_reference.py).
Pinned against the reference's own tokenizer in tests/test_tokenizer_cpu.py (live comparison + committed id vectors).
"""
from __future__ import annotations

import gzip
import html
import os
import re
from typing import Dict, List, Sequence, Tuple, Union

import torch

N_MERGES = 49152 - 256 - 2   # merges used by CLIP (the file holds more lines)
SOT, EOT = "<|startoftext|>", "<|endoftext|>"
PATTERN = r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def default_vocab_path() -> str:
    cands = [os.environ.get("CRIS_BPE_VOCAB"), os.path.join(REPO, "baseline", "_ref", "utils", "bpe_simple_vocab_16e6.txt.gz")]
    for c in cands:
        if c and os.path.exists(c):
            return c
    raise FileNotFoundError("BPE merge table not found: set CRIS_BPE_VOCAB to utils/bpe_simple_vocab_16e6.txt.gz of the "
                            "reference checkout (or stage the reference with tools/stage_reference.py)")


def byte_symbols() -> List[str]:
    """One printable character per byte value: printable Latin-1 bytes stand for themselves, the other 68 are moved
    to code points 256, 257, ... in byte order (so no symbol is whitespace or a control character)."""
    keep = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    out, extra = [], 0
    for b in range(256):
        if b in keep:
            out.append(chr(b))
        else:
            out.append(chr(256 + extra))
            extra += 1
    return out


class BPETokenizer:
    def __init__(self, vocab_path: str = None):
        import regex  # \p{L} / \p{N} classes
        path = vocab_path or default_vocab_path()
        with gzip.open(path, "rt", encoding="utf-8") as f:
            lines = f.read().split("\n")
        pairs = [tuple(l.split()) for l in lines[1:1 + N_MERGES]]
        if len(pairs) != N_MERGES or any(len(p) != 2 for p in pairs):
            raise ValueError(f"{path}: not a CLIP merge table ({len(pairs)} merges)")
        self.rank: Dict[Tuple[str, str], int] = {p: i for i, p in enumerate(pairs)}
        sym = byte_symbols()
        # id order of the reference's vocabulary: byte symbols in ITS table order (printable bytes first, then the moved ones)
        ordered = [s for s in sym if ord(s) < 256] + [s for s in sym if ord(s) >= 256]
        vocab = ordered + [s + "</w>" for s in ordered] + [a + b for a, b in pairs] + [SOT, EOT]
        self.ids: Dict[str, int] = {t: i for i, t in enumerate(vocab)}
        self.words: List[str] = vocab
        self.sym = sym
        self.sot, self.eot = self.ids[SOT], self.ids[EOT]
        self.split = regex.compile(PATTERN, regex.IGNORECASE)
        self._cache: Dict[str, List[int]] = {}
        try:
            import ftfy
            self._fix = ftfy.fix_text
        except ImportError:   # RefCOCO sentences are plain ASCII; without ftfy mojibake repair is skipped
            self._fix = lambda t: t

    def _merge(self, piece: str) -> List[int]:
        hit = self._cache.get(piece)
        if hit is not None:
            return hit
        if piece in (SOT, EOT):
            out = [self.ids[piece]]
        else:
            parts = [self.sym[b] for b in piece.encode("utf-8")]
            parts[-1] += "</w>"
            while len(parts) > 1:
                best, best_rank = None, None
                for a, b in zip(parts, parts[1:]):
                    r = self.rank.get((a, b))
                    if r is not None and (best_rank is None or r < best_rank):
                        best, best_rank = (a, b), r
                if best is None:
                    break
                fused, i = [], 0
                while i < len(parts):
                    if i + 1 < len(parts) and parts[i] == best[0] and parts[i + 1] == best[1]:
                        fused.append(parts[i] + parts[i + 1])
                        i += 2
                    else:
                        fused.append(parts[i])
                        i += 1
                parts = fused
            out = [self.ids[p] for p in parts]
        self._cache[piece] = out
        return out

    def clean(self, text: str) -> str:
        text = html.unescape(html.unescape(self._fix(text))).strip()
        return re.sub(r"\s+", " ", text).strip().lower()

    def encode(self, text: str) -> List[int]:
        out: List[int] = []
        for piece in self.split.findall(self.clean(text)):
            out.extend(self._merge(piece))
        return out

    def decode(self, ids: Sequence[int]) -> str:
        inv = {s: b for b, s in enumerate(self.sym)}
        text = "".join(self.words[int(i)] for i in ids)
        buf = bytearray()   # symbols -> bytes; '</w>' marks the end of a piece
        i = 0
        while i < len(text):
            if text.startswith("</w>", i):
                buf.append(0x20)
                i += 4
            else:
                buf.append(inv[text[i]])
                i += 1
        return buf.decode("utf-8", errors="replace")


_default = None


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False,
             tokenizer: BPETokenizer = None) -> torch.LongTensor:
    """utils/dataset.py:43-84: [SOT] + ids + [EOT] per sentence, zero padded to `context_length`; too long -> cut and the
    last kept id becomes EOT (truncate=True) or RuntimeError (truncate=False).  -> int64 [len(texts), context_length]."""
    global _default
    if tokenizer is None:
        if _default is None:
            _default = BPETokenizer()
        tokenizer = _default
    if isinstance(texts, str):
        texts = [texts]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [tokenizer.sot] + tokenizer.encode(t) + [tokenizer.eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = tokenizer.eot
        out[i, :len(ids)] = torch.tensor(ids)
    return out
