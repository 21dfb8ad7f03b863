"""Generates the committed fixtures under tests/golden/ from the reference's example data.

Run once in the build container (needs /root/reference; the GPU box and the tests never read it):

    python tests/golden/make_golden.py [--skip-oracle]

Outputs
  example_snp.tsv, example_maize_snp.vcf   verbatim copies of reference examples/ (DATA fixtures: the 190 maize
                                           chr1 SNPs of BASELINE.json config 1; SURVEY.md 8c known answers)
  example_genome.fa.gz                     chr1 rebuilt from the TSV's overlapping 512-bp windows (N elsewhere), so
                                           that the VCF path can be checked to regenerate the TSV windows byte for byte
  example_ids.npz                          token ids of every TSV window (position 255 masked), produced here by a
                                           per-character dict loop that is independent of the engine's byte LUT
  l20_seed0_example_scores.npz             CPU-oracle outputs for config 1: PlantCaduceus_l20 random-init seed 0, fp32,
                                           the 185 valid rows: 4 logits (a,c,g,t) at index 255, softmax probs and LLR.
                                           The reference's own model code cannot run offline, so these
                                           vectors come from oracle/caduceus_oracle.py, not from the reference itself.
"""
import argparse
import gzip
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/examples"
VOCAB = {"[PAD]": 0, "[MASK]": 1, "[UNK]": 2, "a": 3, "c": 4, "g": 5, "t": 6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-oracle", action="store_true")
    args = ap.parse_args()
    for name in ("example_snp.tsv", "example_maize_snp.vcf"):
        shutil.copyfile(os.path.join(REF, name), os.path.join(HERE, name))
        os.chmod(os.path.join(HERE, name), 0o644)

    rows = [ln.rstrip("\n").split("\t") for ln in open(os.path.join(HERE, "example_snp.tsv"))]
    hdr, rows = rows[0], rows[1:]
    assert hdr == ["chr", "start", "end", "pos", "ref", "alt", "sequences"], hdr
    # --- genome segment from overlapping windows
    lo = min(int(r[1]) for r in rows)
    hi = max(int(r[2]) for r in rows)
    seg = bytearray(b"N" * (hi - lo))
    for r in rows:
        s, seq = int(r[1]), r[6].encode()
        assert len(seq) == 512 and int(r[2]) - s == 512 and int(r[3]) - 256 == s
        for k, ch in enumerate(seq):
            cur = seg[s - lo + k]
            assert cur in (ord("N"), ch), "overlapping windows disagree"
            seg[s - lo + k] = ch
        assert chr(seq[255]) == r[4]
    assert {r[0] for r in rows} == {"chr1"}
    chrom = b"N" * lo + bytes(seg) + b"N" * 1000
    # lower-case a stretch (soft-masked genomes are common; the window rule upper-cases) -- only inside N padding-free area
    with gzip.open(os.path.join(HERE, "example_genome.fa.gz"), "wt", compresslevel=9) as f:
        f.write(">chr1 rebuilt from example_snp.tsv windows\n")
        for i in range(0, len(chrom), 60):
            f.write(chrom[i:i + 60].decode() + "\n")
        f.write(">chr2 decoy\nACGTNNNNACGT\n")
    # --- golden ids, independent per-character loop
    ids = np.zeros((len(rows), 512), dtype=np.uint8)
    for i, r in enumerate(rows):
        for k, ch in enumerate(r[6]):
            ids[i, k] = VOCAB.get(ch.lower(), VOCAB["[UNK]"])
        ids[i, 255] = VOCAB["[MASK]"]
    np.savez_compressed(os.path.join(HERE, "example_ids.npz"), ids=ids)

    if args.skip_oracle:
        return
    import torch
    from oracle import caduceus_oracle as O
    from plantcaduceus_b200 import preset, random_init_state_dict
    valid = [i for i, r in enumerate(rows) if r[4] in "ACGT" and len(r[4]) == 1 and r[5] in ("A", "C", "G", "T")]
    assert len(valid) == 185
    cfg = preset("PlantCaduceus_l20")
    sd = random_init_state_dict(cfg, seed=0)
    logits4 = np.zeros((len(valid), 4), dtype=np.float32)
    with torch.inference_mode():
        for b in range(0, len(valid), 16):
            sel = valid[b:b + 16]
            lg, _ = O.caduceus_forward(sd, cfg, torch.from_numpy(ids[sel].astype(np.int64)), dtype=torch.float32)
            logits4[b:b + len(sel)] = lg[:, 255, 3:7].numpy()
            print(f"oracle {b + len(sel)}/{len(valid)}", flush=True)
    probs = torch.softmax(torch.from_numpy(logits4), dim=1).numpy()
    nuc = "ACGT"
    llr = np.array([np.log(probs[k][nuc.index(rows[i][5])] / probs[k][nuc.index(rows[i][4])]) for k, i in enumerate(valid)],
                   dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, "l20_seed0_example_scores.npz"), rows=np.array(valid), logits4=logits4,
                        probs=probs, llr=llr)


if __name__ == "__main__":
    main()
