"""VCF + FASTA -> the SNP table ``zero_shot_score.py -input-table`` reads: drop-in for the reference's ``src/format_VCF.sh``

    python -m plantcaduceus_b200.format_vcf <input.vcf> <reference.fasta> <output.tsv>

The shell script (:37-46) pipes every data line through ``awk`` (chr, pos-1, pos, pos, ref, alt), ``bedtools slop -l 255 -r 256``
(interval ``[pos-256, pos+256)`` 0-based, CLIPPED to the chromosome: windows at the ends come out shorter, not N-padded) and
``bedtools getfasta -bedOut -tab`` (the bases as the FASTA spells them, case kept).  samtools / bedtools are not needed
here: the chromosome is sliced directly.  Regenerates the reference's ``examples/example_snp.tsv`` from its
``examples/example_maize_snp.vcf`` byte for byte (tests/test_host_path.py).
"""
from __future__ import annotations

import sys
from typing import Dict, Optional, Sequence, Union

from . import genome_io as gio

HEADER = b"chr\tstart\tend\tpos\tref\talt\tsequences\n"


def format_vcf(vcf_path: str, fasta: Union[str, Dict[str, bytes]], out_path: str, left: int = 255, right: int = 256) -> int:
    """Writes the table; returns the number of rows.  Every VCF record gives one row (multi-allelic ALT strings are kept as
    they are: the scorer drops those rows later, src/zero_shot_score.py:232)."""
    genome = gio.read_fasta(fasta) if isinstance(fasta, str) else fasta
    table = gio.read_vcf_table(vcf_path)
    data, ls, le = table.data, table.line_start.tolist(), table.line_end.tolist()
    out = [HEADER]
    for a, b in zip(ls, le):
        chrom, pos, _id, ref, alt = data[a:b].split(b"\t", 5)[:5]
        name = chrom.decode("utf-8")
        if name not in genome:
            raise KeyError(f"chromosome {name!r} of the VCF is not in the FASTA")
        seq = genome[name]
        p = int(pos)
        start, end = max(0, p - 1 - left), min(len(seq), p + right)
        out.append(b"\t".join((chrom, str(start).encode(), str(end).encode(), pos, ref, alt, seq[start:end])) + b"\n")
    with open(out_path, "wb") as f:
        f.write(b"".join(out))
    return len(out) - 1


def main(argv: Optional[Sequence[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) != 3:
        print("Usage: python -m plantcaduceus_b200.format_vcf <input.vcf> <reference.fasta> <output_file>")
        return 1
    print("Generating contextual sequences...")
    format_vcf(argv[0], argv[1], argv[2])
    print("Done.")
    return 0


if __name__ == "__main__":
    sys.exit(main())
