"""Character-level DNA tokenizer with the subset of the Hugging Face tokenizer API the reference
calls (reference src/zero_shot_score.py:51-57,96,118; src/zero-shot-eval.py:75-126):

    tok.encode_plus(seq, return_tensors="pt", return_attention_mask=False, return_token_type_ids=False)["input_ids"]
    tok(list_of_str, return_tensors="pt", ...)["input_ids"]
    tok.mask_token_id, tok.pad_token_id, tok.get_vocab()["a"|"c"|"g"|"t"], tok.vocab_size

The checkpoint's tokenizer (songlab/tokenizer-dna-mlm: reference pretrain/scripts/train_plant_BERT.py:27-28)
is a lower-casing character tokenizer with no special tokens added, so 512 characters give 512 ids
(reference notebooks/examples.ipynb:132).  Here it is a 256-entry byte LUT applied with numpy (and on the
device by libpcad's tokenize kernel, bit-exact with this table), instead of a per-sequence Python call
in the main process (the reference's host hot loop).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Union

import numpy as np
import torch

from .configuration import DEFAULT_VOCAB


class CharDNATokenizer:
    def __init__(self, vocab: Optional[Dict[str, int]] = None, do_lower_case: bool = True):
        self.vocab = dict(DEFAULT_VOCAB if vocab is None else vocab)
        self.do_lower_case = do_lower_case
        self.pad_token, self.mask_token, self.unk_token = "[PAD]", "[MASK]", "[UNK]"
        for t in (self.pad_token, self.mask_token, self.unk_token):
            if t not in self.vocab:
                raise ValueError(f"vocab lacks {t}")
        lut = np.full(256, self.vocab[self.unk_token], dtype=np.uint8)
        for tok, idx in self.vocab.items():
            if len(tok) == 1:
                lut[ord(tok)] = idx
                if do_lower_case and tok.isalpha():
                    lut[ord(tok.upper())] = idx
        self.lut = lut

    # HF-compatible attributes ------------------------------------------------------------------
    @property
    def mask_token_id(self) -> int:
        return self.vocab[self.mask_token]

    @property
    def pad_token_id(self) -> int:
        return self.vocab[self.pad_token]

    @property
    def unk_token_id(self) -> int:
        return self.vocab[self.unk_token]

    @property
    def vocab_size(self) -> int:
        return len(self.vocab)

    def get_vocab(self) -> Dict[str, int]:
        return dict(self.vocab)

    def __len__(self) -> int:
        return len(self.vocab)

    @classmethod
    def from_pretrained(cls, name_or_path=None, **_kw) -> "CharDNATokenizer":
        """Reads ``vocab.json`` / ``tokenizer.json`` from a local checkpoint directory when present;
        otherwise the documented default vocabulary."""
        import json
        import os
        if name_or_path and os.path.isdir(str(name_or_path)):
            p = os.path.join(str(name_or_path), "tokenizer.json")
            if os.path.exists(p):
                with open(p) as f:
                    tj = json.load(f)
                vocab = tj.get("model", {}).get("vocab")
                if isinstance(vocab, list):         # Unigram-style [[token, score], ...]: the id is the list index
                    vocab = {(e[0] if isinstance(e, (list, tuple)) else e): i for i, e in enumerate(vocab)}
                if isinstance(vocab, dict):
                    vocab = {k: int(v) for k, v in vocab.items()}
                    for t in tj.get("added_tokens") or []:      # special tokens may live only here
                        if isinstance(t, dict) and "content" in t and "id" in t:
                            vocab.setdefault(t["content"], int(t["id"]))
                    return cls(vocab=vocab)
            p = os.path.join(str(name_or_path), "vocab.json")
            if os.path.exists(p):
                with open(p) as f:
                    return cls(vocab={k: int(v) for k, v in json.load(f).items()})
        return cls()

    # encoding ------------------------------------------------------------------------------------
    def encode_bytes(self, seqs: Union[bytes, np.ndarray]) -> np.ndarray:
        """uint8 ASCII -> uint8 ids, any shape."""
        arr = np.frombuffer(seqs, dtype=np.uint8) if isinstance(seqs, (bytes, bytearray)) else np.asarray(seqs, dtype=np.uint8)
        return self.lut[arr]

    def encode(self, seq: str) -> List[int]:
        return self.encode_bytes(seq.encode("latin-1", errors="replace")).tolist()

    def _pack(self, ids, return_tensors):
        if return_tensors == "pt":
            return torch.from_numpy(np.ascontiguousarray(ids).astype(np.int64))
        if return_tensors == "np":
            return np.asarray(ids, dtype=np.int64)
        return ids.tolist()

    def encode_plus(self, text: str, return_tensors: Optional[str] = None, return_attention_mask: bool = False,
                    return_token_type_ids: bool = False, add_special_tokens: bool = False, **_kw):
        ids = self.encode_bytes(text.encode("latin-1", errors="replace"))[None, :]
        out = {"input_ids": self._pack(ids if return_tensors else ids[0], return_tensors)}
        if return_attention_mask:
            out["attention_mask"] = self._pack(np.ones_like(ids if return_tensors else ids[0]), return_tensors)
        if return_token_type_ids:
            out["token_type_ids"] = self._pack(np.zeros_like(ids if return_tensors else ids[0]), return_tensors)
        return out

    def __call__(self, text: Union[str, Sequence[str]], return_tensors: Optional[str] = None,
                 return_attention_mask: bool = False, return_token_type_ids: bool = False,
                 padding: bool = False, add_special_tokens: bool = False, **_kw):
        if isinstance(text, str):
            return self.encode_plus(text, return_tensors=return_tensors, return_attention_mask=return_attention_mask,
                                    return_token_type_ids=return_token_type_ids)
        rows = [self.encode_bytes(t.encode("latin-1", errors="replace")) for t in text]
        lens = {len(r) for r in rows}
        if len(lens) > 1:
            if not padding and return_tensors:
                raise ValueError("sequences of different lengths need padding=True to be returned as a tensor")
            width = max(lens)
            mask = [np.concatenate([np.ones(len(r), np.int64), np.zeros(width - len(r), np.int64)]) for r in rows]
            rows = [np.concatenate([r, np.full(width - len(r), self.pad_token_id, np.uint8)]) for r in rows] if padding else rows
        else:
            mask = [np.ones(len(r), np.int64) for r in rows]
        if return_tensors or len(lens) <= 1 or padding:
            ids = np.stack(rows) if rows else np.zeros((0, 0), np.uint8)
            out = {"input_ids": self._pack(ids, return_tensors)}
            if return_attention_mask:
                out["attention_mask"] = self._pack(np.stack(mask), return_tensors)
        else:
            out = {"input_ids": [r.tolist() for r in rows]}
        return out

    def windows_to_ascii(self, seqs: Iterable[str], length: int) -> np.ndarray:
        """Pack equal-length windows into a uint8 [n, length] ASCII matrix (the engine's host input)."""
        seqs = list(seqs)
        for i, q in enumerate(seqs):
            if len(q) != length:
                raise ValueError(f"window {i} has length {len(q)}, expected {length}")
        flat = "".join(seqs).encode("latin-1", errors="replace")     # one character -> one byte
        return np.frombuffer(flat, dtype=np.uint8).reshape(len(seqs), length).copy()
