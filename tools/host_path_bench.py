"""Host side of the VCF scoring command line, timed on the CPU (no GPU needed): FASTA read, VCF parse, coordinate
extraction and scored-VCF write for N synthetic records, column-wise path (what the CLI runs) beside the per-record path
(the restatement of the reference's PyVCF loop).  Rank 0 does this work serially with the GPUs, so it is sized against the
8-GPU scoring rate (about 7 000 variants/s: DESIGN.md section 6).

    python tools/host_path_bench.py [--records 1000000] [--out profiles/r02_host_path.json]
"""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=1_000_000)
    ap.add_argument("--genome-mb", type=float, default=20.0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import cli_vcf_benchmark as cb
    from plantcaduceus_b200 import genome_io as gio
    from plantcaduceus_b200 import zero_shot_score as zss
    work = tempfile.mkdtemp(prefix="host_path_")
    fa, vcf = os.path.join(work, "g.fa"), os.path.join(work, "v.vcf")
    total = int(args.genome_mb * 1e6)
    chroms = cb.write_genome(fa, [total // 2, total - total // 2])
    cb.write_vcf(vcf, chroms, args.records)
    t = {}

    def timed(name, fn):
        t0 = time.perf_counter()
        out = fn()
        t[name] = round(time.perf_counter() - t0, 3)
        return out

    timed("read_fasta_s", lambda: gio.read_fasta(fa))
    table = timed("read_vcf_table_s", lambda: gio.read_vcf_table(vcf))
    cli_args = zss.parse_args(["-input-vcf", vcf, "-input-fasta", fa, "-output", os.path.join(work, "o.vcf")])
    parsed = timed("variants_from_vcf_s", lambda: zss.variants_from_vcf(cli_args))
    ridx = parsed[4]
    probs = gio.softmax4(np.random.default_rng(0).normal(size=(len(ridx), 4)).astype(np.float32))
    timed("write_scored_vcf_table_s", lambda: gio.write_scored_vcf_table(os.path.join(work, "fast.vcf"), table, ridx, probs))
    header, records = timed("per_record_read_vcf_s", lambda: gio.read_vcf(vcf))
    timed("per_record_write_scored_vcf_s", lambda: gio.write_scored_vcf(os.path.join(work, "slow.vcf"), header, records, list(ridx), probs))
    same = open(os.path.join(work, "fast.vcf"), "rb").read() == open(os.path.join(work, "slow.vcf"), "rb").read()
    line = {"records": args.records, "genome_mb": args.genome_mb, "cpus": os.cpu_count(), "seconds": t,
            "outputs_identical": same,
            "column_wise_parse_plus_write_s_per_million": round((t["read_vcf_table_s"] + t["write_scored_vcf_table_s"]) * 1e6 / args.records, 2),
            "per_record_parse_plus_write_s_per_million": round((t["per_record_read_vcf_s"] + t["per_record_write_scored_vcf_s"]) * 1e6 / args.records, 2)}
    print(json.dumps(line))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(line, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
