"""Tokenizer surface used by the generation runtime.

Real tokenizers (SentencePiece / tiktoken / HF, reference tokenizer.py) are outside the hot path
and need model files that are not available offline; the runtime only needs the interface below.
`SyntheticTokenizer` provides it for synthetic-id runs: ids are their own decoding, and the two id
lists hybrid caches consume (`special_ids`, `punctuation_ids`, reference tokenizer.py:55-63) are
small fixed sets.  `get_tokenizer` returns it when no tokenizer file exists next to the checkpoint.
"""
from __future__ import annotations

import zlib
from pathlib import Path

import torch


class TokenizerInterface:
    def __init__(self, model_path=None):
        self.model_path = model_path

    def encode(self, text):
        raise NotImplementedError

    def decode(self, tokens):
        raise NotImplementedError

    def bos_id(self):
        raise NotImplementedError

    def eos_id(self):
        raise NotImplementedError

    def get_terminator_ids(self):
        return [self.eos_id()]

    def special_ids(self):
        raise NotImplementedError

    def punctuation_ids(self):
        raise NotImplementedError


class SyntheticTokenizer(TokenizerInterface):
    def __init__(self, vocab_size: int = 128256, model_path=None):
        super().__init__(model_path)
        self.vocab_size = vocab_size

    def encode(self, text):
        # a stable hash: the same ids in every process (Python's hash() is salted per process, which would give the
        # ranks of a tensor-parallel run different prompts)
        return [2 + (zlib.crc32(w.encode()) % (self.vocab_size - 2)) for w in str(text).split()]

    def encode_prompt(self, prompt):
        return [self.bos_id()] + self.encode(prompt)

    def decode(self, tokens):
        return " ".join(str(int(t)) for t in tokens)

    def bos_id(self):
        return 0

    def eos_id(self):
        return 1

    def get_terminator_ids(self):
        return None  # synthetic runs never stop early (and never sync on the host)

    def special_ids(self):
        return [[0], [1]]

    def punctuation_ids(self):
        return list(range(2, 34))


def get_tokenizer(tokenizer_model_path, model_name=None, is_chat=False):
    """Synthetic ids when there is no tokenizer file next to the checkpoint.  Real tokenizers (SentencePiece /
    tiktoken / HF) are not part of the hot path and are not re-implemented here: the reference's own tokenizer.py has
    the same interface and is NOT shadowed by dropin/ -- its scripts keep using it."""
    path = Path(str(tokenizer_model_path))
    if path.is_file() or (path.is_dir() and (path / "tokenizer.json").exists()):
        raise NotImplementedError("a tokenizer file was found: use the reference's tokenizer.py (same interface); "
                                  "cold_compress_b200.tokenizer only provides synthetic ids")
    return SyntheticTokenizer()


def encode_tokens(tokenizer, string, bos=True, device="cuda"):
    tokens = tokenizer.encode(string)
    if bos:
        tokens = [tokenizer.bos_id()] + tokens
    return torch.tensor(tokens, dtype=torch.int, device=device)


def encode(tokenizer, prompt, device="cuda", bos=True, is_chat=True):
    """Same signature and meaning as the reference (tokenizer.py:335-349)."""
    if is_chat:
        return torch.tensor(tokenizer.encode_prompt(prompt), dtype=torch.int, device=device)
    return encode_tokens(tokenizer, prompt, device=device, bos=bos)
