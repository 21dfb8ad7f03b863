"""In-silico saturation mutagenesis on the B200 engine (BASELINE.json config 4).

Two modes.

``scan_region`` -- the reference pipeline's semantics.  ``pipelines/in-silico-mutagenesis/1_simulation.R:85-100`` walks a
region, keeps the positions whose base is A/C/G/T and emits one headerless VCF row per (position, alt) -- three per
position, sorted by position (:100-120); ``README.md:56-64`` then feeds those rows to ``src/zero_shot_score.py -input-vcf``,
which gives EVERY row its own 512-bp window CENTRED on the position (``seq_from_vcf``, zero_shot_score.py:187-198) and
runs one forward per row.  Here the chromosome is resident in HBM, the A/C/G/T positions of the region are enumerated
on the device, each position gets its centred window from ``pcad_extract_windows`` and ONE masked forward that serves
its three alts, and the rows come back in the R script's order.  Scores equal the CLI run on the emitted VCF bit for bit
(tests/test_configs_gpu.py).

``saturation_mutagenesis`` -- BASELINE.json config 4 as worded ("all 3 alt alleles at every position of 512bp windows"):
ONE fixed window, every index masked in turn (the context is the same window for every position, NOT a window centred on
it -- a different computation from the pipeline above, kept as the benchmark shape).
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from . import genome_io as gio

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def scan_region(model, chrom, start: int, end: int, batch_size: int = 256, token_idx: int = 255, length: int = 512,
                chrom_dev: Optional[torch.Tensor] = None):
    """Every single-base substitution at every A/C/G/T position of the 1-based inclusive region [start, end] of a
    chromosome (``bytes`` / uint8 array, or already on the device as ``chrom_dev``), each position scored in its own
    window centred at ``token_idx``.  Under torchrun the positions are sharded contiguously over the ranks.

    Returns a dict of arrays with one entry per (position, alt) in the R script's order (position ascending, alt in
    A<C<G<T order, ref skipped): ``pos`` (1-based, int64), ``ref`` / ``alt`` (ASCII codes, uint8), ``score`` =
    log(p_alt / p_ref) (float32) and ``probs`` (float32 [n_positions, 4]) for the positions ``positions`` (1-based)."""
    from . import sharding
    if chrom_dev is None:
        arr = np.frombuffer(chrom, dtype=np.uint8) if isinstance(chrom, (bytes, bytearray)) else np.asarray(chrom, dtype=np.uint8)
        chrom_dev = torch.from_numpy(arr.copy()).to(model.device)
    n_chrom = chrom_dev.numel()
    if start < 1 or end > n_chrom or end < start:
        raise IndexError(f"region [{start}, {end}] outside the chromosome (length {n_chrom})")
    # enumerate on the device: 1_simulation.R keeps `ref %in% c("A","C","G","T")` (soft-masked lower case is dropped)
    region = chrom_dev[start - 1:end]
    acgt = torch.from_numpy(_ACGT.copy()).to(region.device)
    keep = (region[:, None] == acgt[None, :]).any(dim=1)
    pos0_all = torch.nonzero(keep, as_tuple=False).flatten().to(torch.int64) + (start - 1)     # 0-based positions
    n = int(pos0_all.numel())
    rank, _local, world = sharding.env_world()
    lo, hi = sharding.shard_range(n, rank, world)
    logits = torch.empty((hi - lo, 4), dtype=torch.float32, device=model.device)
    for s in range(lo, hi, batch_size):
        e = min(s + batch_size, hi)
        windows = model.extract_windows_device(chrom_dev, pos0_all[s:e], token_idx, length)
        logits[s - lo:e - lo] = model.score_windows_device(windows, token_idx)
    full = sharding.gather_rows(logits, n) if world > 1 else logits
    probs = gio.softmax4(full.cpu().numpy()) if n else np.zeros((0, 4), dtype=np.float32)
    ref = chrom_dev[pos0_all].cpu().numpy()
    ref_idx = np.searchsorted(_ACGT, ref)                      # _ACGT is sorted: A < C < G < T
    alt_idx = np.array([[a for a in range(4) if a != r] for r in range(4)], dtype=np.int64)[ref_idx]   # [n, 3]
    rows = np.repeat(np.arange(n), 3)
    with np.errstate(divide="ignore", invalid="ignore"):
        score = np.log(probs[rows, alt_idx.reshape(-1)] / probs[rows, np.repeat(ref_idx, 3)]).astype(np.float32)
    return {"pos": np.repeat(pos0_all.cpu().numpy() + 1, 3), "ref": np.repeat(ref, 3), "alt": _ACGT[alt_idx.reshape(-1)],
            "score": score, "probs": probs, "positions": pos0_all.cpu().numpy() + 1}


def write_candidate_vcf(path: str, chrom_name: str, result, scores: bool = False) -> None:
    """The headerless 7-column rows ``1_simulation.R:100-120`` writes (chr, pos, '.', ref, alt, '.', '.'), one per
    (position, alt); with ``scores=True`` an eighth INFO column carries ``plantCAD_zero_shot=<score>`` as
    ``zero_shot_score.py -input-vcf`` would add it."""
    pos = np.asarray(result["pos"]).astype(np.int64).astype(str).tolist()
    ref = np.asarray(result["ref"], dtype=np.uint8).tobytes().decode("latin-1")
    alt = np.asarray(result["alt"], dtype=np.uint8).tobytes().decode("latin-1")
    head = chrom_name + "\t"
    if scores:
        sc = np.asarray(result["score"], dtype=np.float32).astype(str).tolist()        # str(np.float32(x)), vectorised
        rows = [f"{head}{p}\t.\t{r}\t{a}\t.\t.\tplantCAD_zero_shot={v}\n" for p, r, a, v in zip(pos, ref, alt, sc)]
    else:
        rows = [f"{head}{p}\t.\t{r}\t{a}\t.\t.\n" for p, r, a in zip(pos, ref, alt)]
    with open(path, "w") as f:
        f.write("".join(rows))


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
