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


_ACGT_CODE = np.full(256, -1, dtype=np.int8)
for _i, _c in enumerate(b"ACGT"):
    _ACGT_CODE[_c] = _i
    _ACGT_CODE[_c + 32] = _i          # lower case


class VcfTable:
    """A VCF held column-wise over the file's own bytes (what ``main`` uses: a 10-million-record VCF parses in seconds
    and costs ~60 bytes of host memory per record beside the file, where one ``VcfRecord`` per line costs minutes and ~1 KB).

    ``data`` is the whole (decompressed) file; record ``i`` is ``data[line_start[i]:line_end[i]]``.  Columns:
    ``chrom_id`` into ``chrom_names`` (order of first appearance), 1-based ``pos``, ``simple`` (1-base REF and a single
    1-base ALT, both in ACGT, any case: the bi-allelic SNP, with ``ref_code`` / ``alt_code`` its indices into ACGT),
    ``has_snv`` (>= 1 SNV ALT: the records the scorer gives a window), ``info_start`` / ``info_end`` (byte range of the
    INFO column, -1 when the line has fewer than 8 columns) and ``n_cols`` (capped at 8).  ``record(i)`` materialises
    the ``VcfRecord`` of one line; ``records()`` all of them (what ``read_vcf`` returns)."""

    def __init__(self, data: bytes, header: List[str], line_start, line_end, chrom_names, chrom_id, pos, ref_code,
                 alt_code, simple, has_snv, info_start, info_end, n_cols):
        self.data, self.header = data, header
        self.line_start, self.line_end = line_start, line_end
        self.chrom_names, self.chrom_id, self.pos = chrom_names, chrom_id, pos
        self.ref_code, self.alt_code, self.simple, self.has_snv = ref_code, alt_code, simple, has_snv
        self.info_start, self.info_end, self.n_cols = info_start, info_end, n_cols

    def __len__(self) -> int:
        return len(self.pos)

    def record(self, i: int) -> VcfRecord:
        fields = self.data[int(self.line_start[i]):int(self.line_end[i])].decode("utf-8").split("\t")
        return VcfRecord(index=int(i), chrom=fields[0], pos=int(fields[1]), ref=fields[3], alts=fields[4].split(","),
                         fields=fields)

    def records(self) -> List[VcfRecord]:
        return [self.record(i) for i in range(len(self))]


def read_vcf_table(path: str) -> VcfTable:
    """Byte-level, vectorised VCF parse (same record set, order and field meaning as ``read_vcf``; reference
    src/zero_shot_score.py:182-201 walks ``vcf.Reader`` one record at a time)."""
    with _open_bytes(path) as f:
        data = f.read()
    arr = np.frombuffer(data, dtype=np.uint8)
    size = arr.size
    if size == 0:
        e64, e8 = np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int8)
        return VcfTable(data, [], e64, e64, [], np.zeros(0, dtype=np.int32), e64, e8, e8, np.zeros(0, dtype=bool),
                        np.zeros(0, dtype=bool), e64, e64, e8)
    nl = np.flatnonzero(arr == 10)
    starts = np.concatenate(([0], nl + 1)).astype(np.int64)
    ends = np.concatenate((nl, [size])).astype(np.int64)
    while True:                                                  # "\r\n" files: drop the carriage returns
        cr = (ends > starts) & (arr[np.maximum(ends - 1, 0)] == 13)
        if not cr.any():
            break
        ends = ends - cr
    keep = ends > starts                                         # blank lines (and the phantom after the last "\n")
    starts, ends = starts[keep], ends[keep]
    is_hdr = arr[starts] == ord("#") if len(starts) else np.zeros(0, dtype=bool)
    header = [data[s:e].decode("utf-8") for s, e in zip(starts[is_hdr].tolist(), ends[is_hdr].tolist())]
    ls, le = starts[~is_hdr], ends[~is_hdr]
    n = len(ls)
    tabs = np.flatnonzero(arr == 9).astype(np.int64)
    first = np.searchsorted(tabs, ls)
    tab_pos, tab_ok = [], []
    for k in range(7):
        j = first + k
        inside = j < len(tabs)
        t = tabs[np.minimum(j, max(len(tabs) - 1, 0))] if len(tabs) else np.zeros(n, dtype=np.int64)
        inside &= t < le
        if k:
            inside &= tab_ok[-1]
        tab_pos.append(np.where(inside, t, le))
        tab_ok.append(inside)
    if n and not tab_ok[3].all():
        bad = int(np.flatnonzero(~tab_ok[3])[0])
        raise ValueError("malformed VCF line (need >= 5 tab-separated columns): "
                         f"{data[int(ls[bad]):int(le[bad])].decode('utf-8', 'replace')[:80]!r}")
    t0, t1, t2, t3, t4, t5, t6 = tab_pos
    n_cols = (1 + sum(ok.astype(np.int8) for ok in tab_ok)).astype(np.int8) if n else np.zeros(0, dtype=np.int8)
    # an 8th column exists iff tab 6 does; INFO ends at the next tab or at the end of the line
    info_start = np.where(tab_ok[6], t6 + 1, -1) if n else np.zeros(0, dtype=np.int64)
    j7 = first + 7
    t7 = tabs[np.minimum(j7, max(len(tabs) - 1, 0))] if len(tabs) else np.zeros(n, dtype=np.int64)
    has7 = tab_ok[6] & (j7 < len(tabs)) & (t7 < le) if n else np.zeros(0, dtype=bool)
    info_end = np.where(tab_ok[6], np.where(has7, t7, le), -1) if n else np.zeros(0, dtype=np.int64)

    # POS: decimal digits between the first two tabs (anything else goes through int(), as in read_vcf)
    w = t1 - t0 - 1
    pos = np.zeros(n, dtype=np.int64)
    odd = w <= 0 if n else np.zeros(0, dtype=bool)
    for k in range(int(min(w.max(), 18)) if n else 0):
        live = k < w
        d = arr[np.minimum(t0 + 1 + k, size - 1)].astype(np.int64) - 48
        odd |= live & ((d < 0) | (d > 9))
        pos = np.where(live, pos * 10 + d, pos)
    if n:
        odd |= w > 18
    for i in np.flatnonzero(odd).tolist():
        pos[i] = int(data[int(t0[i]) + 1:int(t1[i])].decode("utf-8"))

    # CHROM: runs of equal names (a sorted VCF has one run per chromosome)
    chrom_names: List[str] = []
    chrom_id = np.zeros(n, dtype=np.int32)
    if n:
        cw = t0 - ls
        same = cw[1:] == cw[:-1]
        for k in range(int(cw.max())):
            live = (k < cw[1:]) & same
            same &= ~live | (arr[np.minimum(ls[1:] + k, size - 1)] == arr[np.minimum(ls[:-1] + k, size - 1)])
        run_start = np.concatenate(([0], np.flatnonzero(~same) + 1))
        index: Dict[str, int] = {}
        run_ids = np.empty(len(run_start), dtype=np.int32)
        for r, i in enumerate(run_start.tolist()):
            name = data[int(ls[i]):int(t0[i])].decode("utf-8")
            if name not in index:
                index[name] = len(chrom_names)
                chrom_names.append(name)
            run_ids[r] = index[name]
        run_of = np.zeros(n, dtype=np.int64)
        run_of[run_start[1:]] = 1
        chrom_id = run_ids[np.cumsum(run_of)]

    # REF / ALT
    ref_len = t3 - t2 - 1
    alt_len = t4 - t3 - 1
    rc = _ACGT_CODE[arr[np.minimum(t2 + 1, size - 1)]] if n else np.zeros(0, dtype=np.int8)
    ac = _ACGT_CODE[arr[np.minimum(t3 + 1, size - 1)]] if n else np.zeros(0, dtype=np.int8)
    ref_ok = (ref_len == 1) & (rc >= 0)
    simple = ref_ok & (alt_len == 1) & (ac >= 0)
    has_snv = simple.copy()
    for i in np.flatnonzero(ref_ok & (alt_len > 1)).tolist():     # multi-ALT records: any single-base ALT in ACGT
        alts = data[int(t3[i]) + 1:int(t4[i])].split(b",")
        has_snv[i] = any(len(a) == 1 and _ACGT_CODE[a[0]] >= 0 for a in alts)
    ref_code = np.where(simple, rc, -1).astype(np.int8)
    alt_code = np.where(simple, ac, -1).astype(np.int8)
    return VcfTable(data, header, ls, le, chrom_names, chrom_id, pos, ref_code, alt_code, simple, has_snv,
                    info_start, info_end, n_cols)


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
    """softmax over the a,c,g,t logits in float32, computed by the very call the reference makes
    (``torch.nn.functional.softmax(logits.cpu(), dim=1).numpy()``, extract_logits :119), so that identical logits give
    bit-identical probabilities and score strings."""
    import torch
    x = np.ascontiguousarray(logits4, dtype=np.float32)
    if x.size == 0:
        return x.reshape(-1, 4)
    return torch.nn.functional.softmax(torch.from_numpy(x), dim=1).numpy()


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


def write_scored_vcf_table(path: str, table: VcfTable, record_indices, probs: np.ndarray,
                           info_key: str = "plantCAD_zero_shot") -> None:
    """``write_scored_vcf`` over a ``VcfTable``, byte-identical output: the bi-allelic SNP records (``table.simple``) get
    their score from one vectorised float32 log-ratio and their line by splicing the entry into the original bytes;
    only multi-ALT records go through the per-record path."""
    ri = np.asarray(record_indices, dtype=np.int64).reshape(-1)
    probs = np.asarray(probs)
    order = np.argsort(ri, kind="stable")                 # lines leave in record order, like the loop over records
    ri, rows = ri[order], order
    simple = table.simple[ri]
    txt: List[Optional[bytes]] = [None] * len(ri)
    if simple.any():
        k = rows[simple]
        with np.errstate(divide="ignore", invalid="ignore"):
            sc = np.log(probs[k, table.alt_code[ri[simple]].astype(np.int64)] /
                        probs[k, table.ref_code[ri[simple]].astype(np.int64)])
        enc = "\n".join(sc.astype(str).tolist()).encode("ascii").split(b"\n")
        for j, t in zip(np.flatnonzero(simple).tolist(), enc):
            txt[j] = t
    data = table.data
    prefix = (info_key + "=").encode("utf-8")
    out: List[bytes] = []
    declared = any(h.startswith(f"##INFO=<ID={info_key},") for h in table.header)
    for h in table.header:
        if h.startswith("#CHROM") and not declared:
            out.append((f'##INFO=<ID={info_key},Number=A,Type=String,Description="PlantCaduceus zero-shot score '
                        f'log(p_alt/p_ref) per ALT allele; . for non-SNV alleles">\n').encode("utf-8"))
        out.append(h.encode("utf-8") + b"\n")
    cols = (table.line_start[ri].tolist(), table.info_start[ri].tolist(), table.info_end[ri].tolist(),
            table.line_end[ri].tolist(), table.n_cols[ri].tolist(), txt, ri.tolist(), rows.tolist())
    for a, b, c, e, ncol, s, i, k in zip(*cols):
        if s is None:                                     # multi-ALT / non-SNV ALTs: "." per ALT that is not an SNV
            rec = table.record(i)
            p = probs[k]
            ref_p = p[NUCLEOTIDES.index(rec.ref.upper())]
            parts = []
            for alt in rec.alts:
                if rec.alt_is_snv(alt):
                    with np.errstate(divide="ignore", invalid="ignore"):
                        parts.append(str(np.log(p[NUCLEOTIDES.index(alt.upper())] / ref_p)))
                else:
                    parts.append(".")
            s = ",".join(parts).encode("ascii")
        if b < 0:                                         # fewer than 8 columns: pad with "." up to INFO
            out.append(data[a:e] + b"\t." * (7 - ncol) + b"\t" + prefix + s + b"\n")
            continue
        info = data[b:c]
        if info == b"." or info == b"":
            out.append(data[a:b] + prefix + s + data[c:e] + b"\n")
        else:
            out.append(data[a:c] + b";" + prefix + s + data[c:e] + b"\n")
    with open(path, "wb") as f:
        f.write(b"".join(out))
