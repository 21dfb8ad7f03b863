"""This repo's host path and callers against the outputs of the REFERENCE'S OWN CODE executed in the build container
(tests/golden/reference_run/, made by tests/golden/make_reference_run_golden.py: the reference's zero_shot_score.py main(),
seq_from_vcf, zero_shot_score_vcf, zero-shot-eval.py helpers and commands, train_XGBoost.py extract_embeddings -- with the
CPU oracle as the model behind them and data-access stand-ins for PyVCF3 / Biopython).

CPU tests (`not gpu`) put the same oracle behind this repo's callers (tests/oracle_engine.py): everything else -- which rows
are scored, window cutting and padding, tokenisation and masking, which logits are read, softmax / log-ratio arithmetic and
its spelling in the output files, multi-mask row order, SV boundary score, metrics, embedding averaging, the printed
lines -- must then agree with the reference to the last byte or to float32 rounding.  The `gpu` tests run the same
comparisons with the real engine (fp32 parity mode, same seed-0 weights) at the north-star fp32 bar."""
import contextlib
import io
import json
import os

import numpy as np
import pandas as pd
import pytest
import torch

from oracle_engine import OracleEngine, oracle_extract_logits, tiny_config
from plantcaduceus_b200 import CharDNATokenizer
from plantcaduceus_b200 import embeddings as emb_mod
from plantcaduceus_b200 import genome_io as gio
from plantcaduceus_b200 import zero_shot_eval as zse
from plantcaduceus_b200 import zero_shot_score as zss

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RUN = os.path.join(GOLD, "reference_run")


@pytest.fixture(scope="module")
def z():
    return np.load(os.path.join(RUN, "zero_shot_eval.npz"))


@pytest.fixture(scope="module")
def cpu_engine():
    return OracleEngine()


@pytest.fixture(scope="module")
def gpu_engine(cuda_device):
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    cfg, sd = tiny_config()
    return CaduceusForMaskedLM.from_pretrained(sd, config=cfg, torch_dtype=torch.float32).to(cuda_device)


# ---- window rule, record selection ------------------------------------------------------------------------------------
def test_windows_equal_reference_seq_from_vcf():
    g = np.load(os.path.join(RUN, "vcf_windows.npz"))
    fasta = gio.read_fasta(os.path.join(GOLD, "example_genome.fa.gz"))
    for parse in (lambda p: gio.read_vcf(p)[1], lambda p: gio.read_vcf_table(p).records()):
        windows, idx = gio.windows_from_vcf(parse(os.path.join(GOLD, "example_maize_snp.vcf")), fasta, 255, 512)
        assert np.array_equal(np.asarray(idx), g["record_indices"]) and np.array_equal(windows, g["windows"])
    # chromosome ends, soft-masked / N bases, other tokenIdx values
    with open(os.path.join(RUN, "small_windows.json")) as f:
        small = json.load(f)
    fasta = gio.read_fasta(os.path.join(RUN, "small_genome.fa"))
    table = gio.read_vcf_table(os.path.join(RUN, "small.vcf"))
    for tidx, want in small.items():
        windows, idx = gio.windows_from_vcf(table.records(), fasta, int(tidx), 512)
        assert idx == want["record_indices"]
        assert [bytes(w).decode() for w in windows] == want["windows"], f"tokenIdx {tidx}"


# ---- the command line, byte for byte ------------------------------------------------------------------------------------
def _patch_cli(monkeypatch, engine):
    monkeypatch.setattr(zss, "load_model_and_tokenizer", lambda *a, **k: (engine, CharDNATokenizer()))
    monkeypatch.setattr(zss, "extract_logits", oracle_extract_logits)


def test_cli_outputs_equal_reference_main(tmp_path, monkeypatch, cpu_engine):
    _patch_cli(monkeypatch, cpu_engine)
    table = os.path.join(GOLD, "example_snp.tsv")
    assert zss.main(["-input-table", table, "-output", str(tmp_path / "o.tsv"), "-device", "cpu", "-batchSize", "64"]) == 0
    assert zss.main(["-input-table", table, "-output", str(tmp_path / "o.bed"), "-outBED", "-device", "cpu", "-batchSize", "64"]) == 0
    want = pd.read_csv(os.path.join(RUN, "table_scores.tsv"), sep="\t")
    got = pd.read_csv(tmp_path / "o.tsv", sep="\t")
    assert list(got.columns) == list(want.columns) and got.drop(columns="zeroShotScore").equals(want.drop(columns="zeroShotScore"))
    # same model function on both sides; what differs is batch composition inside the oracle's fp32 GEMMs: float32 rounding
    assert np.allclose(got["zeroShotScore"], want["zeroShotScore"], rtol=0, atol=2e-6)
    n_same = sum(a == b for a, b in zip(open(tmp_path / "o.tsv").read().split("\n"), open(os.path.join(RUN, "table_scores.tsv")).read().split("\n")))
    assert n_same == len(want) + 2, n_same                 # header + 185 rows + the empty tail: identical characters
    assert open(tmp_path / "o.tsv", "rb").read() == open(os.path.join(RUN, "table_scores.tsv"), "rb").read()
    assert open(tmp_path / "o.bed", "rb").read() == open(os.path.join(RUN, "table_scores.bed"), "rb").read()


def test_cli_vcf_scores_equal_reference_main(tmp_path, monkeypatch, cpu_engine):
    """INFO/plantCAD_zero_shot of every record as the reference's zero_shot_score_vcf wrote it: which records, one value per
    ALT, '.' for the non-SNV ALTs, str() of a float32."""
    from plantcaduceus_b200 import genome_scan
    _patch_cli(monkeypatch, cpu_engine)
    out = tmp_path / "o.vcf"
    assert zss.main(["-input-vcf", os.path.join(GOLD, "example_maize_snp.vcf"), "-input-fasta", os.path.join(GOLD, "example_genome.fa.gz"),
                     "-output", str(out), "-device", "cpu", "-batchSize", "64"]) == 0
    with open(os.path.join(RUN, "vcf_info.json")) as f:
        want = json.load(f)
    header, recs = gio.read_vcf(str(out))
    assert len(recs) == len(want) == 190
    src = gio.read_vcf(os.path.join(GOLD, "example_maize_snp.vcf"))[1]
    n_exact = 0
    for rec, w in zip(recs, want):
        assert rec.fields[:7] == src[w["record"]].fields[:7]
        got = dict(kv.split("=", 1) for kv in rec.fields[7].split(";") if "=" in kv)["plantCAD_zero_shot"].split(",")
        exp = w["plantCAD_zero_shot"].split(",")
        assert len(got) == len(exp)
        for a, b in zip(got, exp):
            assert (a == ".") == (b == ".")
            if a != ".":
                assert abs(float(a) - float(b)) <= 2e-6
                assert a == str(np.float32(float(a)))    # float32 spelling, not a float64 expansion
                n_exact += a == b
    assert n_exact >= 190, n_exact                       # same logits -> the same characters


# ---- zero-shot-eval.py helpers --------------------------------------------------------------------------------------------
def _check_model_helpers(engine, z, tol):
    tok = CharDNATokenizer()
    seqs = z["seqs"].tolist()
    assert np.abs(zse.masked_probs(engine, tok, seqs, 40, batch_size=3) - z["masked_single_40"]).max() <= tol
    assert np.abs(zse.masked_probs(engine, tok, seqs, [40, 41, 42], batch_size=4) - z["masked_multi_40_41_42"]).max() <= tol
    # rows come back in increasing position order whatever the order of mask_idx (torch.masked_select in the reference)
    assert np.abs(zse.masked_probs(engine, tok, seqs, [70, 5, 41], batch_size=2) - z["masked_multi_unsorted_70_5_41"]).max() <= tol
    assert np.abs(zse.unmasked_probs(engine, tok, seqs, batch_size=3) - z["unmasked"]).max() <= tol
    on_dev = zse.unmasked_probs(engine, tok, seqs, batch_size=3, on_device=True)
    assert np.abs(on_dev.cpu().numpy() - z["unmasked"]).max() <= tol
    g = np.load(os.path.join(RUN, "embeddings.npz"))
    got = emb_mod.extract_embeddings(engine, tok, g["seqs"].tolist(), tokenIdx=int(g["token_idx"]), batch_size=3)
    assert got.shape == g["averaged"].shape == (7, 128)
    assert np.abs(got - g["averaged"]).max() <= tol * max(1.0, np.abs(g["averaged"]).max())


def test_model_helpers_equal_reference_functions_cpu(cpu_engine, z):
    _check_model_helpers(cpu_engine, z, 2e-6)


@pytest.mark.gpu
def test_model_helpers_equal_reference_functions_gpu(gpu_engine, z):
    _check_model_helpers(gpu_engine, z, 1e-4)


def test_metrics_and_sv_score_equal_reference_functions(z):
    assert zse.compute_true_tokens_from_seq(z["seqs"].tolist(), [40]).tolist() == z["true_tokens_40"].tolist()
    assert zse.compute_true_tokens_from_seq(z["seqs"].tolist(), [40, 41, 42]).tolist() == z["true_tokens_40_41_42"].tolist()
    df = pd.DataFrame({"sequence": z["m_seqs"].tolist(), "label": z["m_labels"]})
    assert zse.compute_auroc(df, z["m_probs1"], 5, "sequence") == pytest.approx(float(z["m_auroc_idx5"]), abs=1e-12)
    assert np.array_equal(zse.refprob_scores(df, z["m_probs1"], 5, "sequence"), z["m_refprob_idx5"])
    tt1 = zse.compute_true_tokens_from_seq(df["sequence"], [5])
    tt3 = zse.compute_true_tokens_from_seq(df["sequence"], [4, 5, 6])
    assert zse.metric_token_accuracy(z["m_probs1"], tt1) == pytest.approx(float(z["m_token_acc1"]), abs=1e-12)
    assert zse.metric_token_accuracy(z["m_probs3"], tt3) == pytest.approx(float(z["m_token_acc3"]), abs=1e-12)
    assert zse.metric_motif_accuracy(z["m_probs3"], tt3, 3) == pytest.approx(float(z["m_motif_acc3"]), abs=1e-12)
    assert np.allclose(zse.avg_trueprob_scores(z["m_probs3"], tt3, 3), z["m_avg_trueprob3"], rtol=0, atol=1e-9)
    assert zse.metric_token_accuracy(z["m_probs1"], np.array(["N"] * len(df))) == 0.0
    for fl, key in ((5, "sv_scores_fl5"), (2, "sv_scores_fl2")):
        got = zse.sv_llr_boundary(z["sv_left"], z["sv_right"], z["sv_mut_seqs"].tolist(), z["sv_ref_probs"], z["sv_mut_probs"], fl)
        assert np.allclose(got, z[key], rtol=1e-6, atol=1e-6)
        got_t = zse.sv_llr_boundary(z["sv_left"], z["sv_right"], z["sv_mut_seqs"].tolist(), torch.from_numpy(z["sv_ref_probs"]),
                                    torch.from_numpy(z["sv_mut_probs"]), fl)
        assert np.allclose(got_t, z[key], rtol=1e-5, atol=1e-5)


# ---- the four ZeroShotEval commands end to end ---------------------------------------------------------------------------
def _metric_lines(text):
    return {ln.split("\t")[0]: float(ln.split("\t")[1]) for ln in text.strip().split("\n")}


def _run_commands(tmp_path, monkeypatch, engine, device, tol):
    monkeypatch.setattr(zse, "_load_model", lambda model_name, dev: (engine, CharDNATokenizer()))
    monkeypatch.setattr(zse, "_require_cuda", lambda dev: dev)
    with open(os.path.join(RUN, "eval_runs.json")) as f:
        want = json.load(f)
    ds, sv = os.path.join(RUN, "eval_dataset.tsv"), os.path.join(RUN, "eval_sv_dataset.tsv")

    def run(argv):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            assert zse.main(argv) == 0
        return _metric_lines(buf.getvalue())

    got = run(["evo_cons", "--repo_id", ds, "--task", "t", "--device", device, "--token_idx", "40", "--batch_size=5",
               "--save_logits", str(tmp_path / "evo.tsv"), "--metrics-json", str(tmp_path / "evo.json")])
    exp = _metric_lines(want["evo_cons"]["stdout"])
    assert got.keys() == exp.keys() and all(abs(got[k] - exp[k]) <= tol for k in exp)
    lg, lw = pd.read_csv(tmp_path / "evo.tsv", sep="\t"), pd.read_csv(os.path.join(RUN, "eval_evo_cons_logits.tsv"), sep="\t")
    assert list(lg.columns) == list("ACGT") and np.abs(lg.values - lw.values).max() <= max(tol, 2e-6)
    with open(tmp_path / "evo.json") as f, open(os.path.join(RUN, "eval_evo_cons_metrics.json")) as g:
        a, b = json.load(f), json.load(g)
    assert a.keys() == b.keys() and a["token_idx"] == b["token_idx"] == 40
    got = run(["motif_acc", ds, "t", "--device", device, "--mask-idx=40,41,42", "--motif_len", "3", "--batch_size", "5",
               "--metrics_json", str(tmp_path / "motif.json")])
    exp = _metric_lines(want["motif_acc"]["stdout"])
    assert got.keys() == exp.keys() and all(abs(got[k] - exp[k]) <= tol for k in exp)
    # logits_path: metrics from a saved probability table, no model
    got = run(["motif_acc", ds, "t", "--mask_idx", "(40, 41, 42)", "--logits_path", os.path.join(RUN, "eval_motif_logits.tsv")])
    assert all(abs(got[k] - exp[k]) <= 1e-6 for k in exp)
    got = run(["core_noncore", "--repo_id", ds, "--task", "t", "--device", device, "--mask_idx", "[40,41,42]", "--batch_size", "7"])
    exp = _metric_lines(want["core_noncore"]["stdout"])
    assert got.keys() == exp.keys() and all(abs(got[k] - exp[k]) <= tol for k in exp)
    got = run(["sv_effect", "--repo_id", sv, "--task", "sv", "--device", device, "--batch_size", "5", "--flanking", "5",
               "--output", str(tmp_path / "sv.tsv")])
    exp = _metric_lines(want["sv_effect"]["stdout"])
    assert got.keys() == exp.keys() and all(abs(got[k] - exp[k]) <= tol for k in exp)
    sg, sw = pd.read_csv(tmp_path / "sv.tsv", sep="\t"), pd.read_csv(os.path.join(RUN, "eval_sv_scored.tsv"), sep="\t")
    assert list(sg.columns) == list(sw.columns) and "Left5_Positions" not in sg.columns
    assert np.allclose(sg["score"], sw["score"], rtol=0, atol=max(tol, 1e-5) * 10)


def test_zero_shot_eval_commands_equal_reference_cpu(tmp_path, monkeypatch, cpu_engine):
    _run_commands(tmp_path, monkeypatch, cpu_engine, "cpu", 1e-6)


@pytest.mark.gpu
def test_zero_shot_eval_commands_equal_reference_gpu(tmp_path, monkeypatch, gpu_engine):
    # the metrics are rank statistics of 24 examples: fp32 engine vs fp32 oracle scores differ by ~1e-6, ranks do not move
    _run_commands(tmp_path, monkeypatch, gpu_engine, "cuda:0", 1e-6)


@pytest.mark.gpu
def test_cli_table_scores_equal_reference_main_gpu(tmp_path, monkeypatch, gpu_engine):
    """The real command line (pinned double-buffered batches -> pcad_score_windows_dev) against the reference's main()."""
    monkeypatch.setattr(zss, "load_model_and_tokenizer", lambda *a, **k: (gpu_engine, CharDNATokenizer()))
    assert zss.main(["-input-table", os.path.join(GOLD, "example_snp.tsv"), "-output", str(tmp_path / "o.tsv"), "-batchSize", "64"]) == 0
    want = pd.read_csv(os.path.join(RUN, "table_scores.tsv"), sep="\t")
    got = pd.read_csv(tmp_path / "o.tsv", sep="\t")
    assert got.drop(columns="zeroShotScore").equals(want.drop(columns="zeroShotScore"))
    assert np.abs(got["zeroShotScore"] - want["zeroShotScore"]).max() <= 2e-4      # difference of two fp32 logits at 1e-4 relative


def test_fire_style_argument_parsing():
    assert zse._parse_value("255") == 255 and zse._parse_value("40,41,42") == (40, 41, 42) and zse._parse_value("[1,2]") == [1, 2]
    assert zse._parse_value("cuda:0") == "cuda:0" and zse._parse_value("valid") == "valid"
    assert zse.main([]) == 2 and zse.main(["nope"]) == 2 and zse.main(["evo_cons", "--bogus", "1"]) == 2
    with pytest.raises(RuntimeError, match="CUDA is required"):
        zse._require_cuda("cpu")


# ---- in-silico mutagenesis front end (pipelines/in-silico-mutagenesis/1_simulation.R; R is not installed: literal restatement) ----
def _r_script_rows(chrom: str, regions):
    """1_simulation.R:85-100 spelled out: per region the positions and bases (Biostrings: upper case), rbind, keep A/C/G/T,
    crossing() with the alts (de-duplicates and sorts), drop ref == alt, arrange(pos)."""
    rows = set()
    for s, e in regions:
        for p in range(s, e + 1):
            ref = chrom[p - 1].upper()
            if ref in "ACGT":
                for alt in "ACGT":
                    if alt != ref:
                        rows.add((p, ref, alt))
    return sorted(rows)


def test_mutagenesis_front_end_matches_r_script_semantics(tmp_path, monkeypatch, cpu_engine):
    from plantcaduceus_b200 import mutagenesis as mut
    rng = np.random.default_rng(23)
    chrom = "".join(rng.choice(list("ACGTacgtNRY"), size=900, p=[0.2, 0.2, 0.2, 0.2, 0.04, 0.04, 0.04, 0.04, 0.02, 0.01, 0.01]))
    other = "ACGT" * 50
    (tmp_path / "g.fa").write_text(">chrA x\n" + "\n".join(chrom[i:i + 70] for i in range(0, 900, 70)) + "\n>chrB\n" + other + "\n")
    gff = ["##gff-version 3", "chrA\tsrc\tgene\t120\t140\t.\t+\t.\tID=g1", "chrA\tsrc\tmRNA\t120\t140\t.\t+\t.\tID=m1;Parent=g1",
           "chrA\tsrc\texon\t125\t135\t.\t+\t.\tParent=m1", "chrA\tsrc\tgene\t135\t160\t.\t-\t.\tID=g2",      # overlaps g1 once extended
           "chrA\tsrc\tgene\t5\t30\t.\t+\t.\tID=g3",                                                           # 5 - 10 < 1: dropped
           "chrA\tsrc\tgene\t880\t895\t.\t+\t.\tID=g4",                                                        # 895 + 10 > 900: dropped
           "chrB\tsrc\tgene\t50\t60\t.\t+\t.\tID=g5", "chrA\tsrc\tgene\t600\t610\t.\t+\t.\tID=g6", "##FASTA", ">chrA", "ACGT"]
    (tmp_path / "a.gff").write_text("\n".join(gff) + "\n")
    regions = mut.gene_regions_from_gff(str(tmp_path / "a.gff"), "chrA", flank=10, chrom_len=900)
    assert regions == [(110, 150), (125, 170), (590, 620)]
    assert mut.merge_regions(regions) == [(110, 170), (590, 620)]
    want = _r_script_rows(chrom, regions)
    got = mut.enumerate_candidates(chrom.encode(), regions)
    assert list(zip(got["pos"].tolist(), map(chr, got["ref"]), map(chr, got["alt"]))) == want
    # the command line without --score == the R script's output file
    assert mut.main(["-g", str(tmp_path / "a.gff"), "-f", str(tmp_path / "g.fa"), "-o", str(tmp_path / "cand.vcf"), "-c", "chrA", "-k", "10"]) == 0
    assert (tmp_path / "cand.vcf").read_text() == "".join(f"chrA\t{p}\t.\t{r}\t{a}\t.\t.\n" for p, r, a in want)
    assert mut.main(["-g", str(tmp_path / "a.gff"), "-f", str(tmp_path / "g.fa"), "-o", str(tmp_path / "x.vcf"), "-c", "chrZ"]) == 1
    # scored in the same run == the candidate file pushed through the scoring command line (README.md:56-64 of the pipeline)
    res = mut.scan_regions(cpu_engine, chrom.encode(), regions, batch_size=64, length=512)
    assert list(zip(res["pos"].tolist(), map(chr, res["ref"]), map(chr, res["alt"]))) == want
    _patch_cli(monkeypatch, cpu_engine)
    assert zss.main(["-input-vcf", str(tmp_path / "cand.vcf"), "-input-fasta", str(tmp_path / "g.fa"), "-output", str(tmp_path / "s.vcf"),
                     "-device", "cpu", "-batchSize", "64"]) == 0
    rows = [ln.rstrip("\n").split("\t") for ln in open(tmp_path / "s.vcf") if not ln.startswith("#")]
    assert len(rows) == len(want)
    cli = np.array([np.float32(r[7].split("plantCAD_zero_shot=")[1]) for r in rows])
    assert np.allclose(cli, res["score"], rtol=0, atol=3e-6)          # one forward per row there, per position here (batching)
    mut.write_candidate_vcf(str(tmp_path / "scored.vcf"), "chrA", res, scores=True)
    assert sum(1 for _ in open(tmp_path / "scored.vcf")) == len(want)


def test_config1_golden_is_what_the_reference_main_writes():
    """BASELINE.json configs[0]: the committed oracle vectors (tests/golden/l20_seed0_example_scores.npz, which the GPU command
    line is compared with in tests/test_scoring_cli_gpu.py) equal, as float32, the zeroShotScore column the reference's own
    main() wrote for the same model (reference_run/config1_scores.tsv, `make_reference_run_golden.py --config1`)."""
    g = np.load(os.path.join(GOLD, "l20_seed0_example_scores.npz"))
    df = pd.read_csv(os.path.join(RUN, "config1_scores.tsv"), sep="\t")
    src = pd.read_csv(os.path.join(GOLD, "example_snp.tsv"), sep="\t").iloc[g["rows"]]
    assert len(df) == 185 and df["pos"].tolist() == src["pos"].tolist() and df["alt"].tolist() == src["alt"].tolist()
    assert np.array_equal(df["zeroShotScore"].to_numpy().astype(np.float32), g["llr"].astype(np.float32))
