"""Model-facing helpers of the reference's ``src/zero-shot-eval.py`` on the B200 engine:
masked-position probabilities for single / multi-mask inputs (`_masked_probs`, :129-140), per-position
probabilities for whole windows (`_unmasked_probs`, :143-178) and the structural-variant boundary score
(`_sv_llr_boundary`, :181-243).  The benchmark datasets (HF ``datasets``) and the sklearn metrics around them are
host post-processing and stay with the caller.

Differences in how the work is done: sequences travel as ASCII bytes and are tokenised on the device; for masked
inputs the LM head runs only at the masked positions (``pcad_score_masked``) instead of producing ``[B, L, V]``
logits and selecting afterwards; the boundary score is vectorised numpy instead of a per-row Python loop.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import genome_io as gio

NUCLEOTIDES = ("A", "C", "G", "T")


def _ascii_matrix(tokenizer, sequences: Sequence[str]) -> np.ndarray:
    seqs = [str(s) for s in sequences]
    if not seqs:
        return np.zeros((0, 0), dtype=np.uint8)
    L = len(seqs[0])
    for s in seqs:
        if len(s) != L:
            raise ValueError(f"All sequences must have same length; got {len(s)} vs {L}")
    return tokenizer.windows_to_ascii(seqs, L)


def masked_probs(model, tokenizer, sequences: Sequence[str], mask_idx: Sequence[int], batch_size: int = 64) -> np.ndarray:
    """Softmax over a,c,g,t at every masked position: float32 ``[N * len(mask_idx), 4]``, sequence-major then in
    increasing position order -- the order ``torch.masked_select`` yields in the reference (:136-137).
    ``mask_idx`` may be one index (SingleMaskDataset) or several (MultiMaskDataset)."""
    mask_idx = sorted(int(i) for i in ([mask_idx] if np.isscalar(mask_idx) else mask_idx))
    ascii_mat = _ascii_matrix(tokenizer, sequences)
    n, L = ascii_mat.shape
    if n and mask_idx and mask_idx[-1] >= L:
        raise AssertionError("mask index out of range")
    out = np.zeros((n * len(mask_idx), 4), dtype=np.float32)
    pos_row = np.asarray(mask_idx, dtype=np.int32)
    for s in range(0, n, batch_size):
        ids = tokenizer.encode_bytes(ascii_mat[s:s + batch_size])
        ids[:, pos_row] = tokenizer.mask_token_id
        pos = np.repeat(pos_row[None, :], len(ids), axis=0)
        logits4 = model.score_masked(torch.from_numpy(ids), torch.from_numpy(pos)).cpu().numpy()
        out[s * len(mask_idx):(s + len(ids)) * len(mask_idx)] = gio.softmax4(logits4.reshape(-1, 4))
    return out


def unmasked_probs(model, tokenizer, sequences: Sequence[str], batch_size: int = 16, on_device: bool = False):
    """Per-position probabilities over A,C,G,T for each (unmasked) sequence: float32 ``[N, L, 4]`` (:143-178).
    ``on_device=True`` returns a torch tensor that stays on the model's device (at L = 8192 the ``[N, L, 4]`` array is what
    crosses PCIe otherwise; ``sv_llr_boundary`` accepts it and gathers its 2 * flanking values per row there)."""
    ascii_mat = _ascii_matrix(tokenizer, sequences)
    n, L = ascii_mat.shape
    v = tokenizer.get_vocab()
    idxs = [v[c] for c in "acgt"]
    if on_device:
        out_dev = torch.zeros((n, L, 4), dtype=torch.float32, device=model.device)
        sel = torch.tensor(idxs, device=model.device)
        for s in range(0, n, batch_size):
            ids = torch.from_numpy(tokenizer.encode_bytes(ascii_mat[s:s + batch_size]).astype(np.int64))
            logits = model(input_ids=ids.to(model.device)).logits
            out_dev[s:s + len(ids)] = torch.softmax(logits.index_select(-1, sel).float(), dim=-1)
        return out_dev
    out = np.zeros((n, L, 4), dtype=np.float32)
    for s in range(0, n, batch_size):
        ids = torch.from_numpy(tokenizer.encode_bytes(ascii_mat[s:s + batch_size]).astype(np.int64))
        logits = model(input_ids=ids.to(model.device)).logits.cpu().numpy()[..., idxs]
        out[s:s + len(ids)] = gio.softmax4(logits.reshape(-1, 4)).reshape(len(ids), L, 4)
    return out


def sv_llr_boundary(left: Sequence[int], right: Sequence[int], mut_seqs: Sequence[str], ref_probs: np.ndarray,
                    mut_probs: np.ndarray, flanking: int) -> np.ndarray:
    """-mean log(p_mut / p_ref) over the 2*flanking boundary positions of each structural variant (:181-243).
    ``left`` / ``right`` are the 1-based breakpoints in the reference window; the mutated window is centred at
    ``L // 2``.  Reference positions: ``[left-flanking .. left-1]`` and ``[right+1 .. right+flanking]`` (1-based);
    mutated positions: ``[c-flanking .. c-1]`` and ``[c .. c+flanking-1]`` (0-based).  The channel is the mutated base
    at the mutated position; non-ACGT bases contribute 0; probabilities are floored at 1e-12."""
    n = len(mut_seqs)
    L = ref_probs.shape[1]
    c = L // 2
    if isinstance(ref_probs, torch.Tensor) or isinstance(mut_probs, torch.Tensor):
        return _sv_llr_boundary_device(left, right, mut_seqs, ref_probs, mut_probs, flanking)
    k = np.arange(flanking)
    left = np.asarray(left, dtype=np.int64)
    right = np.asarray(right, dtype=np.int64)
    # 0-based reference positions, [n, 2*flanking] in (left block, right block) order
    ref_pos = np.concatenate([(left[:, None] - 1) - (flanking - 1) + k[None, :] - 1, (right[:, None] + 1) + k[None, :] - 1], axis=1)
    mut_pos = np.concatenate([c - flanking + k, c + k])[None, :].repeat(n, axis=0)
    centre = np.array([list(str(s)[c - flanking:c + flanking].upper().ljust(2 * flanking, "N")) for s in mut_seqs])
    chan = np.full(centre.shape, -1, dtype=np.int64)
    for j, b in enumerate(NUCLEOTIDES):
        chan[centre == b] = j
    valid = chan >= 0
    rows = np.arange(n)[:, None].repeat(2 * flanking, axis=1)
    ch = np.where(valid, chan, 0)
    r = ref_probs[rows, ref_pos, ch]
    m = mut_probs[rows, mut_pos, ch]
    vals = np.where(valid, np.log(np.maximum(m, 1e-12) / np.maximum(r, 1e-12)), 0.0)
    return -vals.mean(axis=1)


def _sv_llr_boundary_device(left, right, mut_seqs, ref_probs, mut_probs, flanking: int) -> np.ndarray:
    """``sv_llr_boundary`` with the probability tensors resident on the device: the 2 * flanking (position, channel) pairs of
    every row are gathered there and only the ``[n]`` scores come back."""
    dev = ref_probs.device if isinstance(ref_probs, torch.Tensor) else mut_probs.device
    ref_probs = torch.as_tensor(ref_probs, device=dev)
    mut_probs = torch.as_tensor(mut_probs, device=dev)
    n, L = len(mut_seqs), ref_probs.shape[1]
    c = L // 2
    k = torch.arange(flanking, device=dev)
    left_t = torch.as_tensor(np.asarray(left, dtype=np.int64), device=dev)
    right_t = torch.as_tensor(np.asarray(right, dtype=np.int64), device=dev)
    ref_pos = torch.cat([(left_t[:, None] - 1) - (flanking - 1) + k[None, :] - 1, (right_t[:, None] + 1) + k[None, :] - 1], dim=1)
    mut_pos = torch.cat([c - flanking + k, c + k])[None, :].expand(n, -1)
    centre = np.array([list(str(s)[c - flanking:c + flanking].upper().ljust(2 * flanking, "N")) for s in mut_seqs])
    chan = np.full(centre.shape, -1, dtype=np.int64)
    for j, b in enumerate(NUCLEOTIDES):
        chan[centre == b] = j
    chan_t = torch.as_tensor(chan, device=dev)
    valid = chan_t >= 0
    ch = torch.where(valid, chan_t, torch.zeros_like(chan_t))
    rows = torch.arange(n, device=dev)[:, None].expand(-1, 2 * flanking)
    r = ref_probs[rows, ref_pos, ch]
    m = mut_probs[rows, mut_pos, ch]
    vals = torch.where(valid, torch.log(torch.clamp(m, min=1e-12) / torch.clamp(r, min=1e-12)), torch.zeros_like(m))
    return (-vals.mean(dim=1)).cpu().numpy()


def compute_true_tokens_from_seq(sequences: Sequence[str], positions: List[int]) -> np.ndarray:
    return np.array([str(seq)[i].upper() for seq in sequences for i in positions])
