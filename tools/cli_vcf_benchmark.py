"""BASELINE.json config 3 through the real command line: ``python -m plantcaduceus_b200.zero_shot_score -input-vcf ...
-input-fasta ...`` on a synthetic genome and a synthetic sorted VCF, timed by wall clock around a fresh process, at two
variant counts.  The difference quotient (T_large - T_small) / (n_large - n_small) is the CLI's steady-state cost per
variant with everything the reference's ``seq_from_vcf`` / ``zero_shot_score_vcf`` path does per record (parse, window,
tokenise, score, write: src/zero_shot_score.py:137-214) and none of the fixed start-up; the intercept is the start-up
(interpreter, torch import, CUDA context, weight upload, FASTA read).  Compare the slope with bench.py's device-resident
rate: the host side must not be the limit.

    python tools/cli_vcf_benchmark.py --out gpurun_out/r02_cli_vcf_benchmark.json
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def write_genome(path: str, lengths, seed: int = 2):
    rng = np.random.default_rng(seed)
    chroms = {}
    with open(path, "wb") as f:
        for k, n in enumerate(lengths):
            codes = rng.integers(0, 4, size=n, dtype=np.uint8)
            chroms[f"chr{k + 1}"] = codes
            seq = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
            pad = (-n) % 60
            lines = np.concatenate([seq, np.full(pad, ord("A"), np.uint8)]).reshape(-1, 60)
            body = np.concatenate([lines, np.full((len(lines), 1), 10, np.uint8)], axis=1).tobytes()
            if pad:                                  # drop the padding of the last line again
                body = body[:-(pad + 1)] + b"\n"
            f.write(f">chr{k + 1} synthetic\n".encode() + body)
    return chroms


def write_vcf(path: str, chroms, n: int, seed: int = 3):
    """n sorted records spread over the chromosomes in proportion to their length; 2 % carry a second ALT (an insertion:
    written as '.' by the scorer), REF is the genome's base."""
    rng = np.random.default_rng(seed)
    total = sum(len(c) for c in chroms.values())
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        left = n
        names = list(chroms)
        for k, name in enumerate(names):
            codes = chroms[name]
            m = left if k == len(names) - 1 else int(round(n * len(codes) / total))
            left -= m
            pos = np.sort(rng.choice(len(codes), size=m, replace=False))
            ref = codes[pos]
            alt = (ref + rng.integers(1, 4, size=m)) % 4
            multi = rng.random(m) < 0.02
            rows = [f"{name}\t{p + 1}\t.\t{'ACGT'[r]}\t{'ACGT'[a]}{',' + 'ACGT'[r] + 'TG' if mu else ''}\t.\tPASS\t.\n"
                    for p, r, a, mu in zip(pos.tolist(), ref.tolist(), alt.tolist(), multi.tolist())]
            f.write("".join(rows))


def run_cli(vcf, fasta, out, model, batch, device):
    cmd = [sys.executable, "-m", "plantcaduceus_b200.zero_shot_score", "-input-vcf", vcf, "-input-fasta", fasta, "-output", out,
           "-model", model, "-device", device, "-batchSize", str(batch)]
    t0 = time.perf_counter()
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    dt = time.perf_counter() - t0
    if p.returncode != 0:
        print(p.stdout[-3000:], file=sys.stderr)
        raise SystemExit(1)
    with open(out) as f:
        rows = sum(1 for ln in f if not ln.startswith("#"))
    return dt, rows


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="PlantCaduceus_l32")
    ap.add_argument("--n-small", type=int, default=6000)
    ap.add_argument("--n-large", type=int, default=36000)
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--genome-mb", type=float, default=30.0)
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    work = tempfile.mkdtemp(prefix="cli_vcf_bench_")
    total = int(args.genome_mb * 1e6)
    fasta = os.path.join(work, "genome.fa")
    chroms = write_genome(fasta, [total // 2, total // 3, total - total // 2 - total // 3])
    runs = []
    for label, n in (("warmup", 256), ("small", args.n_small), ("large", args.n_large)):
        vcf = os.path.join(work, f"{label}.vcf")
        write_vcf(vcf, chroms, n)
        dt, rows = run_cli(vcf, fasta, os.path.join(work, f"{label}.out.vcf"), args.model, args.batch_size, args.device)
        assert rows == n, (rows, n)
        runs.append({"run": label, "variants": n, "seconds": round(dt, 2)})
        print(json.dumps(runs[-1]), flush=True)
    small, large = runs[1], runs[2]
    slope = (large["seconds"] - small["seconds"]) / (large["variants"] - small["variants"])
    line = {"benchmark": "config 3 through the CLI (-input-vcf / -input-fasta), wall clock of a fresh process", "model": args.model,
            "batch_size": args.batch_size, "genome_mb": args.genome_mb, "hardware": "1 x B200", "dtype": "bf16",
            "weights": "random-init preset", "runs": runs, "steady_state_variants_per_s": round(1.0 / slope, 1),
            "startup_seconds": round(small["seconds"] - slope * small["variants"], 2)}
    print(json.dumps(line))
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        with open(args.out, "w") as f:
            json.dump(line, f, indent=1)
    return 0


if __name__ == "__main__":
    sys.exit(main())
