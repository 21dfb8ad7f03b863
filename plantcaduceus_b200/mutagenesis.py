"""In-silico saturation mutagenesis on the B200 engine (BASELINE.json config 4).

``scan_region`` / ``scan_regions`` -- the reference pipeline's semantics.  ``pipelines/in-silico-mutagenesis/1_simulation.R``
takes the ``gene`` features of one chromosome from a GFF, extends them by ``--flank`` on both sides, drops the ones that
leave the chromosome (:68-79), walks every position of every region, keeps the positions whose base is A/C/G/T, crosses
them with the four alts and drops ref == alt (:85-100) -- ``tidyr::crossing`` de-duplicates and sorts, so the result is the
UNION of the regions, one row per (position, alt), position ascending, alt in A<C<G<T order -- and writes headerless
7-column VCF rows (:106-122).  The genome goes through Biostrings (``readDNAStringSet`` -> 2bit -> ``getSeq``), which
has no lower case: soft-masked bases ARE mutated, with an upper-case ref; IUPAC ambiguity codes become N and are dropped.
``README.md`` of that pipeline then feeds the rows to ``src/zero_shot_score.py -input-vcf``, which gives EVERY row its own
512-bp window CENTRED on the position (``seq_from_vcf``, zero_shot_score.py:187-198) and runs one forward per row.  Here the
chromosome is resident in HBM, the A/C/G/T positions of the region are enumerated on the device, each position gets its
centred window from ``pcad_extract_windows`` and ONE masked forward that serves its three alts, and the rows come back in
the R script's order.  Scores equal the CLI run on the emitted VCF bit for bit (tests/test_configs_gpu.py).

    python -m plantcaduceus_b200.mutagenesis -g genes.gff -f genome.fa -o candidates.vcf -c chr1 [-k 2000] [--score -model <dir|preset>]

writes what ``Rscript 1_simulation.R`` writes (and, with ``--score``, an eighth column ``plantCAD_zero_shot=<score>``).

``saturation_mutagenesis`` -- BASELINE.json config 4 as worded ("all 3 alt alleles at every position of 512bp windows"):
ONE fixed window, every index masked in turn (the context is the same window for every position, NOT a window centred on
it -- a different computation from the pipeline above, kept as the benchmark shape).
"""
from __future__ import annotations

import sys
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import genome_io as gio

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_ALT_IDX = np.array([[a for a in range(4) if a != r] for r in range(4)], dtype=np.int64)      # [ref, 3] alts in A<C<G<T order


def gene_regions_from_gff(path: str, chrom: str, flank: int = 2000, chrom_len: Optional[int] = None) -> List[Tuple[int, int]]:
    """1-based inclusive ``[start - flank, end + flank]`` of every ``gene`` feature of ``chrom`` in a GFF/GFF3 file, in file
    order, without the ones that would start before base 1 or end after ``chrom_len`` (1_simulation.R:68-79:
    ``resize(width + 2 * flank, fix = "center")`` moves both ends by exactly ``flank``)."""
    regions = []
    with gio._open_text(path) as f:
        for line in f:
            if line.startswith("##FASTA"):
                break
            if not line.strip() or line.startswith("#"):
                continue
            c = line.rstrip("\n").split("\t")
            if len(c) < 5 or c[2] != "gene" or c[0] != chrom:
                continue
            s, e = int(c[3]) - flank, int(c[4]) + flank
            if s > 0 and (chrom_len is None or e <= chrom_len):
                regions.append((s, e))
    return regions


def merge_regions(regions: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """Sorted union of 1-based inclusive intervals (overlapping or touching ones fused): the position set
    ``crossing()`` leaves after de-duplication (1_simulation.R:97)."""
    out: List[List[int]] = []
    for s, e in sorted((int(s), int(e)) for s, e in regions):
        if out and s <= out[-1][1] + 1:
            out[-1][1] = max(out[-1][1], e)
        else:
            out.append([s, e])
    return [(s, e) for s, e in out]


def enumerate_candidates(chrom, regions: Sequence[Tuple[int, int]]):
    """The rows 1_simulation.R emits for ``regions`` of one chromosome, on the host and without a model: dict of ``pos``
    (1-based), ``ref``, ``alt`` (ASCII codes), three rows per A/C/G/T position (any case; ref upper-cased)."""
    arr = np.frombuffer(chrom, dtype=np.uint8) if isinstance(chrom, (bytes, bytearray)) else np.asarray(chrom, dtype=np.uint8)
    pos_parts = []
    for s, e in merge_regions(regions):
        if s < 1 or e > len(arr):
            raise IndexError(f"region [{s}, {e}] outside the chromosome (length {len(arr)})")
        seg = arr[s - 1:e]
        up = np.where((seg >= 97) & (seg <= 122), seg - 32, seg)
        pos_parts.append(np.flatnonzero(np.isin(up, _ACGT)) + (s - 1))
    pos0 = np.concatenate(pos_parts) if pos_parts else np.zeros(0, dtype=np.int64)
    ref = arr[pos0]
    ref = np.where((ref >= 97) & (ref <= 122), ref - 32, ref).astype(np.uint8)
    alt_idx = _ALT_IDX[np.searchsorted(_ACGT, ref)]
    return {"pos": np.repeat(pos0 + 1, 3), "ref": np.repeat(ref, 3), "alt": _ACGT[alt_idx.reshape(-1)]}


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
    # enumerate on the device: 1_simulation.R keeps `ref %in% c("A","C","G","T")` of a genome that went through Biostrings,
    # i.e. upper-cased: soft-masked bases count, N and IUPAC codes do not
    region = chrom_dev[start - 1:end]
    region = torch.where((region >= 97) & (region <= 122), region - 32, region)
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
    ref = np.where((ref >= 97) & (ref <= 122), ref - 32, ref).astype(np.uint8)
    ref_idx = np.searchsorted(_ACGT, ref)                      # _ACGT is sorted: A < C < G < T
    alt_idx = _ALT_IDX[ref_idx]                                # [n, 3]
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


def scan_regions(model, chrom, regions: Sequence[Tuple[int, int]], batch_size: int = 256, token_idx: int = 255, length: int = 512,
                 chrom_dev: Optional[torch.Tensor] = None):
    """``scan_region`` over the union of several regions of one chromosome (the extended gene regions of 1_simulation.R):
    every position once, rows in position order.  Same dict as ``scan_region``."""
    if chrom_dev is None:
        arr = np.frombuffer(chrom, dtype=np.uint8) if isinstance(chrom, (bytes, bytearray)) else np.asarray(chrom, dtype=np.uint8)
        chrom_dev = torch.from_numpy(arr.copy()).to(model.device)
    parts = [scan_region(model, None, s, e, batch_size, token_idx, length, chrom_dev=chrom_dev) for s, e in merge_regions(regions)]
    if not parts:
        return {"pos": np.zeros(0, dtype=np.int64), "ref": np.zeros(0, dtype=np.uint8), "alt": np.zeros(0, dtype=np.uint8),
                "score": np.zeros(0, dtype=np.float32), "probs": np.zeros((0, 4), dtype=np.float32), "positions": np.zeros(0, dtype=np.int64)}
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}


def main(argv: Optional[Sequence[str]] = None) -> int:
    """``Rscript 1_simulation.R -g GFF -f FASTA -o OUTPUT -c CHR [-k FLANK]`` (:21-33), optionally scoring the rows in the
    same run (``--score``) instead of a second pass through ``zero_shot_score.py -input-vcf``."""
    import argparse
    ap = argparse.ArgumentParser(description="Simulate SNPs in extended gene regions from a GFF and FASTA file.")
    ap.add_argument("-g", "--gff", required=True, help="Path to the input GFF file (e.g., annotations.gff).")
    ap.add_argument("-f", "--fasta", required=True, help="Path to the input genome FASTA file (e.g., genome.fa).")
    ap.add_argument("-o", "--output", required=True, help="Path for the output file (e.g., potential_snps.vcf).")
    ap.add_argument("-c", "--chr", required=True, help="Target chromosome name (e.g., 'chr1'). Must match names in GFF/FASTA.")
    ap.add_argument("-k", "--flank", type=int, default=2000, help="Flank size in base pairs to extend gene regions on both sides")
    ap.add_argument("--score", action="store_true", help="also score every row (adds INFO plantCAD_zero_shot=<log(p_alt/p_ref)>)")
    ap.add_argument("-model", "--model", default="PlantCaduceus_l32")
    ap.add_argument("-device", "--device", default="cuda:0")
    ap.add_argument("-batchSize", "--batch-size", dest="batch_size", type=int, default=256)
    args = ap.parse_args(argv)
    fasta = gio.read_fasta(args.fasta)
    if args.chr not in fasta:
        print(f"Error: Chromosome '{args.chr}' not found in the FASTA file. Please check your chromosome names.", file=sys.stderr)
        return 1
    chrom = fasta[args.chr]
    regions = gene_regions_from_gff(args.gff, args.chr, args.flank, len(chrom))
    if args.score:
        from .zero_shot_score import load_model_and_tokenizer
        model, _tok = load_model_and_tokenizer(args.model, args.device)
        result = scan_regions(model, chrom, regions, batch_size=args.batch_size)
    else:
        result = enumerate_candidates(chrom, regions)
    if len(result["pos"]) == 0:
        print("Warning: No candidate SNPs were generated. The output file will be empty.")
    write_candidate_vcf(args.output, args.chr, result, scores=args.score)
    return 0


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


if __name__ == "__main__":
    sys.exit(main())
