"""In-silico saturation mutagenesis on the B200 engine (BASELINE.json config 4).

The reference enumerates every single-base substitution with an R script (pipelines/in-silico-mutagenesis/
1_simulation.R:85-100: all 3 alternative alleles at every A/C/G/T position) and then feeds the resulting VCF
rows to ``src/zero_shot_score.py``, i.e. one masked forward per *position* serves its 3 variants
(zero_shot_score.py:154-160).  Here the enumeration is index arithmetic: window w becomes ``len(positions)`` rows
that differ only in which index is masked, scored in batches with ``pcad_score_masked``.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import genome_io as gio

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def saturation_mutagenesis(model, window: str, positions: Optional[Sequence[int]] = None, batch_size: int = 256):
    """Scores every substitution at ``positions`` (default: every A/C/G/T position) of one window.

    Returns a dict of arrays, one entry per (position, alt) pair in position-major, A<C<G<T order:
    ``pos`` (0-based index in the window), ``ref``, ``alt`` (ASCII codes), ``score`` = log(p_alt / p_ref) with the
    probabilities taken from the forward pass that masks that position.
    """
    tok = model._tokenizer
    ascii_row = np.frombuffer(window.encode("latin-1"), dtype=np.uint8)
    upper = np.frombuffer(window.upper().encode("latin-1"), dtype=np.uint8)
    L = len(ascii_row)
    if positions is None:
        positions = np.nonzero(np.isin(upper, _ACGT))[0]
    positions = np.asarray(positions, dtype=np.int64)
    if positions.size and (positions.min() < 0 or positions.max() >= L):
        raise IndexError("position outside the window")
    ids_row = tok.encode_bytes(ascii_row)
    probs = np.zeros((len(positions), 4), dtype=np.float32)
    for s in range(0, len(positions), batch_size):
        p = positions[s:s + batch_size]
        ids = np.repeat(ids_row[None, :], len(p), axis=0)
        ids[np.arange(len(p)), p] = tok.mask_token_id
        logits4 = model.score_masked(torch.from_numpy(ids), torch.from_numpy(p.astype(np.int32))[:, None])
        probs[s:s + len(p)] = gio.softmax4(logits4[:, 0].cpu().numpy())
    out_pos, out_ref, out_alt, out_score = [], [], [], []
    for k, pos in enumerate(positions):
        ref = upper[pos]
        if ref not in _ACGT:
            continue   # the R generator only mutates A/C/G/T positions
        r = int(np.nonzero(_ACGT == ref)[0][0])
        for a in range(4):
            if a == r:
                continue
            out_pos.append(int(pos)); out_ref.append(int(ref)); out_alt.append(int(_ACGT[a]))
            with np.errstate(divide="ignore", invalid="ignore"):
                out_score.append(float(np.log(probs[k, a] / probs[k, r])))
    return {"pos": np.array(out_pos, dtype=np.int64), "ref": np.array(out_ref, dtype=np.uint8),
            "alt": np.array(out_alt, dtype=np.uint8), "score": np.array(out_score, dtype=np.float32)}
