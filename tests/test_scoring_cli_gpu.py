"""BASELINE.json config 1 end to end on the GPU: the zero_shot_score-compatible CLI (plantcaduceus_b200/zero_shot_score.py
-> CaduceusForMaskedLM.score_windows_host -> libpcad C ABI) on the reference's example table / VCF, compared with the
committed CPU-oracle golden vectors (tests/golden/l20_seed0_example_scores.npz; PlantCaduceus_l20, random-init seed 0)."""
import os

import numpy as np
import pytest

from plantcaduceus_b200 import genome_io as gio
from plantcaduceus_b200 import zero_shot_score as zss

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def spearman(a, b):
    ra = np.argsort(np.argsort(a)).astype(np.float64)
    rb = np.argsort(np.argsort(b)).astype(np.float64)
    return float(np.corrcoef(ra, rb)[0, 1])


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "l20_seed0_example_scores.npz"))


@pytest.mark.parametrize("dtype,tol", [("float32", None), ("bfloat16", 2e-2)])
def test_table_scores_match_golden(cuda_device, tmp_path, gold, dtype, tol):
    import pandas as pd
    out = tmp_path / f"scores_{dtype}.tsv"
    rc = zss.main(["-input-table", os.path.join(GOLD, "example_snp.tsv"), "-output", str(out), "-model", "PlantCaduceus_l20",
                   "-device", "cuda:0", "-batchSize", "64", "-dtype", dtype, "-seed", "0"])
    assert rc == 0
    df = pd.read_csv(out, sep="\t")
    assert list(df.columns) == ["chr", "start", "end", "pos", "ref", "alt", "sequences", "zeroShotScore"]
    assert len(df) == 185
    got = df["zeroShotScore"].to_numpy(dtype=np.float64)
    want = gold["llr"].astype(np.float64)
    if dtype == "float32":
        # fp32 logits within 1e-4 relative => LLR (a difference of two logits) within 2e-4 of the logit scale
        scale = np.abs(gold["logits4"]).max()
        assert np.abs(got - want).max() <= 2e-4 * scale
    else:
        assert np.abs(got - want).max() <= 2 * tol      # two bf16 logits, each within 2e-2 absolute
    assert spearman(got, want) >= 0.999


def test_vcf_and_bed_outputs(cuda_device, tmp_path, gold):
    import pandas as pd
    out_vcf = tmp_path / "scored.vcf"
    rc = zss.main(["-input-vcf", os.path.join(GOLD, "example_maize_snp.vcf"), "-input-fasta",
                   os.path.join(GOLD, "example_genome.fa.gz"), "-output", str(out_vcf), "-model", "PlantCaduceus_l20",
                   "-dtype", "float32", "-batchSize", "128"])
    assert rc == 0
    header, recs = gio.read_vcf(str(out_vcf))
    assert len(recs) == 190
    rows = {int(r): k for k, r in enumerate(gold["rows"])}
    scale = np.abs(gold["logits4"]).max()
    n_checked = 0
    for i, rec in enumerate(recs):
        info = dict(kv.split("=", 1) for kv in rec.fields[7].split(";") if "=" in kv)
        vals = info["plantCAD_zero_shot"].split(",")
        assert len(vals) == len(rec.alts)
        if i in rows:            # bi-allelic SNP: same variant as the table row
            assert abs(float(vals[0]) - float(gold["llr"][rows[i]])) <= 2e-4 * scale
            n_checked += 1
        else:                    # multi-allelic: "." exactly for the non-SNV ALTs
            assert [v == "." for v in vals] == [not rec.alt_is_snv(a) for a in rec.alts]
    assert n_checked == 185
    out_bed = tmp_path / "scores.bed"
    rc = zss.main(["-input-table", os.path.join(GOLD, "example_snp.tsv"), "-output", str(out_bed), "-outBED", "-model",
                   "PlantCaduceus_l20", "-dtype", "float32"])
    assert rc == 0
    bed = pd.read_csv(out_bed, sep="\t", header=None)
    assert bed.shape == (185, 6)
    assert (bed[2] - bed[1] == 1).all()
