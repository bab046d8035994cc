"""CLIP byte-pair-encoding tokenizer for real prompt strings.

The reference never names it: diffusers loads ``CLIPTokenizer`` with the pipeline (run_aug/run_aug.py:202-211) and the filter calls
``clip.tokenize`` (all_utils/utils.py:301,310).  Both are the same algorithm (openai-clip ``clip/simple_tokenizer.py``, restated here
from its published description): lower-case + whitespace-normalise the text, split it with CLIP's pattern, map every UTF-8 byte of a
piece to a printable code point (GPT-2's byte table), mark the last symbol of the piece with ``</w>`` and greedily apply the ranked
merges; ids come from a vocabulary of 256 byte symbols, their 256 ``</w>`` forms, one entry per merge, and the two specials.

The vocabulary files are NOT available offline (``tokenizer/vocab.json`` + ``merges.txt`` of an HF repo, or openai-clip's
``bpe_simple_vocab_16e6.txt.gz``): ``from_files`` / ``from_openai_bpe`` read them when a checkpoint directory is present; tests build a
synthetic merges table and compare against ``transformers.CLIPTokenizer`` constructed from the very same files.

Two call conventions, as the two callers use them:
  * ``tok(texts)``            -- HF ``padding="max_length", truncation=True``: BOS + ids[:L-2] + EOS, padded with ``pad_id`` (SD v1.5 pads
                                 with EOS, SDXL's tokenizer_2 with "!" = id 0);
  * ``tok.tokenize(texts)``   -- ``clip.tokenize``: zero padding, RuntimeError when the text does not fit (unless ``truncate``).
"""
from __future__ import annotations

import gzip
import html
import json
from functools import lru_cache
from typing import Dict, List, Optional, Sequence, Tuple, Union

import torch

BOS, EOS = "<|startoftext|>", "<|endoftext|>"


@lru_cache()
def bytes_to_unicode() -> Dict[int, str]:
    """GPT-2's reversible byte -> printable code point table: the 188 printable Latin-1 bytes map to themselves, the other 68 to
    U+0100.. in byte order."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + list(range(ord("\xae"), ord("\xff") + 1))
    table, extra = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def _pattern():
    import regex

    return regex.compile(r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""", regex.IGNORECASE)


def base_vocab(merges: Sequence[Tuple[str, str]]) -> Dict[str, int]:
    """The id assignment of openai-clip: byte symbols, byte symbols + '</w>', merges in rank order, then the two specials."""
    symbols = list(bytes_to_unicode().values())
    vocab = symbols + [s + "</w>" for s in symbols] + ["".join(m) for m in merges] + [BOS, EOS]
    return {tok: i for i, tok in enumerate(vocab)}


class CLIPBPETokenizer:
    def __init__(self, encoder: Dict[str, int], merges: Sequence[Tuple[str, str]], model_max_length: int = 77, pad_id: Optional[int] = None):
        self.encoder = dict(encoder)
        self.ranks = {tuple(m): i for i, m in enumerate(merges)}
        self.model_max_length = model_max_length
        self.bos_id, self.eos_id = self.encoder[BOS], self.encoder[EOS]
        self.pad_id = self.eos_id if pad_id is None else pad_id
        self.vocab_size = len(self.encoder)
        self.byte_encoder = bytes_to_unicode()
        self.pat = _pattern()
        self._cache: Dict[str, Tuple[str, ...]] = {BOS: (BOS,), EOS: (EOS,)}

    # ---- construction ---------------------------------------------------------------------------------------------------
    @staticmethod
    def _read_merges(lines: Sequence[str], limit: Optional[int] = None) -> List[Tuple[str, str]]:
        lines = [ln for ln in lines[1:] if ln.strip()]  # the first line is a version banner
        if limit is not None:
            lines = lines[:limit]
        return [tuple(ln.split()) for ln in lines]

    @classmethod
    def from_files(cls, vocab_json, merges_txt, pad_token: Optional[str] = None, model_max_length: int = 77) -> "CLIPBPETokenizer":
        """HF layout: tokenizer/vocab.json (token -> id) + tokenizer/merges.txt."""
        encoder = json.load(open(vocab_json, encoding="utf-8"))
        merges = cls._read_merges(open(merges_txt, encoding="utf-8").read().split("\n"), 49152 - 256 - 2)
        pad_id = encoder[pad_token] if pad_token is not None else None
        return cls(encoder, merges, model_max_length, pad_id)

    @classmethod
    def from_openai_bpe(cls, bpe_path, model_max_length: int = 77) -> "CLIPBPETokenizer":
        """openai-clip's bpe_simple_vocab_16e6.txt.gz: the vocabulary is DERIVED from the first 48894 merges."""
        merges = cls._read_merges(gzip.open(bpe_path).read().decode("utf-8").split("\n"), 49152 - 256 - 2)
        return cls(base_vocab(merges), merges, model_max_length, pad_id=0)

    @classmethod
    def from_merges(cls, merges: Sequence[Tuple[str, str]], **kw) -> "CLIPBPETokenizer":
        return cls(base_vocab(merges), merges, **kw)

    # ---- algorithm ----------------------------------------------------------------------------------------------------------
    @staticmethod
    def clean(text: str) -> str:
        """basic_clean + whitespace_clean + lower (ftfy's mojibake repair, which openai-clip applies first, is not available offline;
        it is the identity on well-formed text)."""
        text = html.unescape(html.unescape(text))
        return " ".join(text.split()).strip().lower()

    def bpe(self, token: str) -> Tuple[str, ...]:
        if token in self._cache:
            return self._cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = {(word[i], word[i + 1]) for i in range(len(word) - 1)}
            best = min(pairs, key=lambda p: self.ranks.get(p, float("inf")))
            if best not in self.ranks:
                break
            first, second = best
            merged, i = [], 0
            while i < len(word):
                if i < len(word) - 1 and word[i] == first and word[i + 1] == second:
                    merged.append(first + second)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = tuple(merged)
        self._cache[token] = word
        return word

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        for piece in self.pat.findall(self.clean(text)):
            piece = "".join(self.byte_encoder[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[t] for t in self.bpe(piece))
        return ids

    def decode(self, ids: Sequence[int]) -> str:
        inv = {v: k for k, v in self.encoder.items()}
        byte_dec = {v: k for k, v in self.byte_encoder.items()}
        text = "".join(inv[int(i)] for i in ids)
        words = text.replace(BOS, "").replace(EOS, "").split("</w>")
        return " ".join(bytearray(byte_dec[c] for c in w).decode("utf-8", errors="replace") for w in words if w)

    # ---- the two call conventions -----------------------------------------------------------------------------------------------
    def __call__(self, texts: Union[str, Sequence[str]], max_length: Optional[int] = None) -> torch.Tensor:
        L = max_length or self.model_max_length
        if isinstance(texts, str):
            texts = [texts]
        rows = []
        for t in texts:
            ids = [self.bos_id] + self.encode(t)[: L - 2] + [self.eos_id]
            rows.append(ids + [self.pad_id] * (L - len(ids)))
        return torch.tensor(rows, dtype=torch.int64)

    def tokenize(self, texts: Union[str, Sequence[str]], context_length: int = 77, truncate: bool = False) -> torch.Tensor:
        if isinstance(texts, str):
            texts = [texts]
        out = torch.zeros((len(texts), context_length), dtype=torch.int64)
        for r, t in enumerate(texts):
            ids = [self.bos_id] + self.encode(t) + [self.eos_id]
            if len(ids) > context_length:
                if not truncate:
                    raise RuntimeError(f"Input {t} is too long for context length {context_length}")
                ids = ids[:context_length]
                ids[-1] = self.eos_id
            out[r, : len(ids)] = torch.tensor(ids)
        return out
