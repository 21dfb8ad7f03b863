"""Host-side data formats either side of the scoring path: FASTA genome, VCF records, SNP tables.

Replaces, for this path only, what the reference gets from Biopython (``SeqIO.to_dict``), PyVCF3
(``vcf.Reader`` / ``vcf.Writer``) and pandas in ``src/zero_shot_score.py`` (:137-214, :228-258); none of
those packages is needed.  Everything here is byte-oriented: a chromosome is one ``bytes`` object and a
batch of windows is a ``uint8 [n, 512]`` matrix that goes to the GPU as is (tokenisation happens there).
"""
from __future__ import annotations

import gzip
import io
from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

NUCLEOTIDES = ("A", "C", "G", "T")


def _open_text(path: str):
    return gzip.open(path, "rt") if str(path).endswith(".gz") else open(path, "rt")


# ---------------------------------------------------------------------------------------------------
# FASTA
# ---------------------------------------------------------------------------------------------------
def _open_bytes(path: str):
    return gzip.open(path, "rb") if str(path).endswith(".gz") else open(path, "rb")


_WHITESPACE = np.zeros(256, dtype=bool)
_WHITESPACE[[9, 10, 11, 12, 13, 32]] = True


def read_fasta(path: str) -> Dict[str, bytes]:
    """{record id: sequence bytes}.  The id is the header up to the first whitespace, which is the key
    ``SeqIO.to_dict`` uses (reference src/zero_shot_score.py:176-180).  The file is read once as bytes and each record's
    line breaks are removed with one vectorised mask (a per-line Python loop is minutes on a plant genome)."""
    with _open_bytes(path) as f:
        data = f.read()
    arr = np.frombuffer(data, dtype=np.uint8)
    out: Dict[str, bytes] = {}
    if arr.size == 0:
        return out
    nl = np.flatnonzero(arr == 10)
    starts = np.concatenate(([0], nl + 1))
    starts = starts[starts < arr.size]
    hdr = starts[arr[starts] == ord(">")]
    for k, h in enumerate(hdr):
        j = np.searchsorted(nl, h)
        eol = int(nl[j]) if j < len(nl) else arr.size
        fields = data[h + 1:eol].split()
        name = fields[0].decode("ascii") if fields else ""
        if name in out:
            raise ValueError(f"duplicate FASTA record id {name!r}")
        body = arr[min(eol + 1, arr.size):int(hdr[k + 1]) if k + 1 < len(hdr) else arr.size]
        out[name] = body[~_WHITESPACE[body]].tobytes()
    return out


# ---------------------------------------------------------------------------------------------------
# VCF
# ---------------------------------------------------------------------------------------------------
@dataclass
class VcfRecord:
    index: int              # 0-based index among data lines (the reference's recordIdx)
    chrom: str
    pos: int                # 1-based
    ref: str
    alts: List[str]
    fields: List[str]       # the raw tab-split line, for writing back

    def alt_is_snv(self, alt: str) -> bool:
        """PyVCF's ``alt.type == "SNV"`` for a substitution: a single-base REF replaced by a single base.
        Symbolic (<DEL>), breakend, '*' and '.' alleles and any length change are not SNVs."""
        return len(self.ref) == 1 and len(alt) == 1 and alt.upper() in NUCLEOTIDES and self.ref.upper() in NUCLEOTIDES

    @property
    def has_snv(self) -> bool:
        return any(self.alt_is_snv(a) for a in self.alts)


def read_vcf(path: str) -> Tuple[List[str], List[VcfRecord]]:
    """Returns (header lines incl. the #CHROM line, records)."""
    header: List[str] = []
    records: List[VcfRecord] = []
    with _open_text(path) as f:
        for line in f:
            line = line.rstrip("\n").rstrip("\r")
            if not line:
                continue
            if line.startswith("#"):
                header.append(line)
                continue
            fields = line.split("\t")
            if len(fields) < 5:
                raise ValueError(f"malformed VCF line (need >= 5 tab-separated columns): {line[:80]!r}")
            records.append(VcfRecord(index=len(records), chrom=fields[0], pos=int(fields[1]), ref=fields[3],
                                     alts=fields[4].split(","), fields=fields))
    return header, records


# ---------------------------------------------------------------------------------------------------
# windows
# ---------------------------------------------------------------------------------------------------
def extract_window(chrom_seq: bytes, pos0: int, token_idx: int, length: int = 512) -> bytes:
    """The 512-bp context of a variant at 0-based ``pos0`` with the variant at index ``token_idx``:
    ``chrom[pos0-token_idx : pos0+length-token_idx]`` upper-cased; padded with ``N`` on the left when the
    window would start before the chromosome (right-justified) and on the right otherwise
    (reference src/zero_shot_score.py:185-198; same interval as src/format_VCF.sh:42-44)."""
    add = length - token_idx
    if pos0 - token_idx < 0:
        s = chrom_seq[0:max(pos0 + add, 0)].upper()
        return s.rjust(length, b"N")
    s = chrom_seq[pos0 - token_idx:pos0 + add].upper()
    return s.ljust(length, b"N")


def windows_from_vcf(records: Sequence[VcfRecord], fasta: Dict[str, bytes], token_idx: int = 255,
                     length: int = 512) -> Tuple[np.ndarray, List[int]]:
    """One window per record that has at least one SNV ALT (one forward pass serves all its ALTs).
    Returns (uint8 ASCII [n, length], record indices)."""
    rows: List[bytes] = []
    idx: List[int] = []
    for rec in records:
        if not rec.has_snv:
            continue
        if rec.chrom not in fasta:
            raise KeyError(f"VCF record {rec.index}: chromosome {rec.chrom!r} is not in the FASTA "
                           f"(check that chromosome names match)")
        rows.append(extract_window(fasta[rec.chrom], rec.pos - 1, token_idx, length))
        idx.append(rec.index)
    if not rows:
        return np.zeros((0, length), dtype=np.uint8), idx
    return np.frombuffer(b"".join(rows), dtype=np.uint8).reshape(len(rows), length).copy(), idx


# ---------------------------------------------------------------------------------------------------
# scores
# ---------------------------------------------------------------------------------------------------
def softmax4(logits4: np.ndarray) -> np.ndarray:
    """softmax over the a,c,g,t logits in float32 (reference extract_logits, :119)."""
    x = np.asarray(logits4, dtype=np.float32)
    x = x - x.max(axis=1, keepdims=True)
    e = np.exp(x)
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def llr(probs: np.ndarray, ref_idx: np.ndarray, alt_idx: np.ndarray) -> np.ndarray:
    """log(p_alt / p_ref) per row, vectorised (reference zero_shot_score, :124-134)."""
    rows = np.arange(len(probs))
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.log(probs[rows, alt_idx] / probs[rows, ref_idx])


def write_scored_vcf(path: str, header: Sequence[str], records: Sequence[VcfRecord], record_indices: Sequence[int],
                     probs: np.ndarray, info_key: str = "plantCAD_zero_shot") -> None:
    """Writes the records that were scored, each with ``INFO/plantCAD_zero_shot`` = comma-joined
    log(p_alt/p_ref) per ALT, ``.`` for non-SNV ALTs (reference zero_shot_score_vcf, :137-169).
    Unlike PyVCF's writer this also declares the INFO key in the header."""
    by_index = {ri: k for k, ri in enumerate(record_indices)}
    buf = io.StringIO()
    declared = any(h.startswith(f"##INFO=<ID={info_key},") for h in header)
    for h in header:
        if h.startswith("#CHROM") and not declared:
            buf.write(f'##INFO=<ID={info_key},Number=A,Type=String,Description="PlantCaduceus zero-shot score '
                      f'log(p_alt/p_ref) per ALT allele; . for non-SNV alleles">\n')
        buf.write(h + "\n")
    for rec in records:
        k = by_index.get(rec.index)
        if k is None:
            continue
        p = probs[k]
        ref_p = p[NUCLEOTIDES.index(rec.ref.upper())]
        scores = []
        for alt in rec.alts:
            if rec.alt_is_snv(alt):
                with np.errstate(divide="ignore", invalid="ignore"):
                    scores.append(str(np.log(p[NUCLEOTIDES.index(alt.upper())] / ref_p)))
            else:
                scores.append(".")
        fields = list(rec.fields)
        while len(fields) < 8:
            fields.append(".")
        entry = f"{info_key}=" + ",".join(scores)
        fields[7] = entry if fields[7] in (".", "") else fields[7] + ";" + entry
        buf.write("\t".join(fields) + "\n")
    with open(path, "w") as f:
        f.write(buf.getvalue())
