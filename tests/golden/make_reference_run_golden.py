"""Golden vectors from the REFERENCE'S OWN CODE, run here in the build container.

Everything of the scoring path that lives in the reference repository itself (as opposed to the un-vendored model
packages) is imported from /root/reference and EXECUTED, and its outputs are committed under tests/golden/reference_run/:

  src/zero_shot_score.py   main() on examples/example_snp.tsv (TSV and -outBED), and on examples/example_maize_snp.vcf with the
                           genome of tests/golden/: the filter, SequenceDataset (tokenise + mask), the DataLoader batching,
                           extract_logits, zero_shot_score, seq_from_vcf's window rule, zero_shot_score_vcf's per-ALT scores
  src/zero-shot-eval.py    SingleMaskDataset / MultiMaskDataset + _masked_probs, _unmasked_probs, _sv_llr_boundary and the
                           metric helpers (_compute_auroc, _refprob_scores, _metric_token_accuracy, _metric_motif_accuracy,
                           _avg_trueprob_scores, _compute_true_tokens_from_seq)
  src/train_XGBoost.py     SequenceDataset + extract_embeddings

What is NOT the reference's: (1) the MODEL those functions call -- the hub model code and mamba_ssm are absent, so
``load_model_and_tokenizer`` is replaced by the CPU oracle (oracle/caduceus_oracle.py, fp32, a 2-layer d_model 128
random-init seed-0 Caduceus) behind the HF call surface, with this repo's CharDNATokenizer; (2) duck-typed stand-ins for
the absent I/O packages, holding NO logic of the path: ``vcf`` (PyVCF3: Reader yields records with CHROM / POS / REF / ALT
/ INFO, ALT alleles typed the way PyVCF's ``_Substitution`` types them: "SNV" iff one character; the Writer only records
what main() put into INFO), ``Bio.SeqIO`` (parse / to_dict / record slicing), and empty ``fire`` / ``xgboost`` /
``matplotlib`` modules that are imported but never called here.

So these vectors pin, against the reference's own statements: which rows are scored, how windows are cut and padded, how
they are tokenised and masked, which logits are read, the softmax / log-ratio arithmetic and its output spelling, the row
order of multi-mask probabilities, the SV boundary score, the metrics, the embedding averaging.  They do NOT pin the model
arithmetic.

    python tests/golden/make_reference_run_golden.py        # needs /root/reference; tests and the GPU box never do
"""
import gzip
import importlib.util
import json
import os
import sys
import types
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(HERE, "reference_run")


# ---------------------------------------------------------------------------------------------------------------------
# stand-ins for absent packages (data access only)
# ---------------------------------------------------------------------------------------------------------------------
class _Substitution:
    """PyVCF3 model.py: a plain ALT string is a substitution, typed "SNV" iff it is one character long, else "MNV"."""

    def __init__(self, nucleotides):
        self.sequence = str(nucleotides)
        self.type = "SNV" if len(self.sequence) == 1 else "MNV"

    def __str__(self):
        return self.sequence


class _SV:
    def __init__(self, text):
        self.type = text.strip("<>")
        self.text = text

    def __str__(self):
        return self.text


class _Record:
    def __init__(self, fields):
        self.fields = fields
        self.CHROM, self.POS, self.ID, self.REF = fields[0], int(fields[1]), fields[2], fields[3]
        self.ALT = [(_SV(a) if a.startswith("<") else (None if a == "." else _Substitution(a))) for a in fields[4].split(",")]
        info = OrderedDict()
        if len(fields) > 7 and fields[7] not in (".", ""):
            for kv in fields[7].split(";"):
                k, _, v = kv.partition("=")
                info[k] = v if _ else True
        self.INFO = info

    def __str__(self):
        return f"Record(CHROM={self.CHROM}, POS={self.POS}, REF={self.REF}, ALT={[str(a) for a in self.ALT]})"


WRITTEN = []          # (record index among data lines, INFO["plantCAD_zero_shot"]) in the order main() wrote them


def _install_stubs():
    vcf = types.ModuleType("vcf")

    class Reader:
        def __init__(self, fsock=None, filename=None, **_kw):
            self.filename = filename
            opener = gzip.open if str(filename).endswith(".gz") else open
            with opener(filename, "rt") as f:
                self._lines = [ln.rstrip("\n") for ln in f if ln.strip() and not ln.startswith("#")]

        def __iter__(self):
            for k, ln in enumerate(self._lines):
                rec = _Record(ln.split("\t"))
                rec._index = k
                yield rec

    class Writer:
        def __init__(self, stream, template, **_kw):
            self.stream = stream

        def write_record(self, record):
            WRITTEN.append((record._index, record.INFO["plantCAD_zero_shot"]))

        def close(self):
            self.stream.close()

    vcf.Reader, vcf.Writer = Reader, Writer
    sys.modules["vcf"] = vcf

    class SeqRecord:
        def __init__(self, rid, seq):
            self.id, self.seq = rid, seq            # ``str(record.seq)`` is all the reference asks of it

        def __getitem__(self, sl):
            return SeqRecord(self.id, self.seq[sl])

    def parse(handle, fmt):
        assert fmt == "fasta"
        close = False
        if isinstance(handle, str):
            handle, close = open(handle, "rt"), True
        rid, chunks = None, []
        for ln in handle:
            ln = ln.rstrip("\r\n")
            if ln.startswith(">"):
                if rid is not None:
                    yield SeqRecord(rid, "".join(chunks))
                rid, chunks = ln[1:].split()[0], []
            elif rid is not None:
                chunks.append(ln.strip())
        if rid is not None:
            yield SeqRecord(rid, "".join(chunks))
        if close:
            handle.close()

    def to_dict(records):
        out = {}
        for r in records:
            if r.id in out:
                raise ValueError(f"Duplicate key '{r.id}'")
            out[r.id] = r
        return out

    bio = types.ModuleType("Bio")
    seqio = types.ModuleType("Bio.SeqIO")
    seqio.parse, seqio.to_dict = parse, to_dict
    bio.SeqIO = seqio
    sys.modules["Bio"], sys.modules["Bio.SeqIO"] = bio, seqio
    def absent(top):
        try:
            return importlib.util.find_spec(top) is None
        except (ValueError, ImportError):
            return True

    for top, names in (("fire", ["fire"]), ("xgboost", ["xgboost"]), ("matplotlib", ["matplotlib", "matplotlib.pyplot"])):
        if absent(top):
            for name in names:
                sys.modules[name] = types.ModuleType(name)
            if len(names) > 1:
                sys.modules[names[0]].pyplot = sys.modules[names[1]]


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ---------------------------------------------------------------------------------------------------------------------
# the model behind the HF call surface: the CPU oracle
# ---------------------------------------------------------------------------------------------------------------------
def tiny_model():
    from oracle import caduceus_oracle as O
    from plantcaduceus_b200 import CaduceusConfig, CharDNATokenizer, random_init_state_dict
    cfg = CaduceusConfig(d_model=128, n_layer=2)
    sd = random_init_state_dict(cfg, seed=0)

    class OracleModel:
        config = cfg

        def to(self, *a, **k):
            return self

        def eval(self):
            return self

        def __call__(self, input_ids=None, output_hidden_states=False, **_kw):
            with torch.inference_mode():
                logits, hs = O.caduceus_forward(sd, cfg, input_ids, dtype=torch.float32, output_hidden_states=output_hidden_states)
            return SimpleNamespace(logits=logits, hidden_states=tuple(hs) if hs is not None else None)

    return OracleModel(), CharDNATokenizer()


def config1(zss):
    """BASELINE.json configs[0] through the reference's main(): PlantCaduceus_l20 random-init seed 0 (the oracle, fp32) on
    examples/example_snp.tsv -> config1_scores.tsv (about seven minutes of CPU)."""
    from oracle import caduceus_oracle as O
    from plantcaduceus_b200 import CharDNATokenizer, preset, random_init_state_dict
    cfg = preset("PlantCaduceus_l20")
    sd = random_init_state_dict(cfg, seed=0)

    class L20:
        def to(self, *a, **k):
            return self

        def eval(self):
            return self

        def __call__(self, input_ids=None, **_kw):
            with torch.inference_mode():
                logits, _ = O.caduceus_forward(sd, cfg, input_ids, dtype=torch.float32)
            return SimpleNamespace(logits=logits)

    zss.load_model_and_tokenizer = lambda model_dir, device: (L20(), CharDNATokenizer())
    sys.argv = ["zero_shot_score.py", "-input-table", os.path.join(HERE, "example_snp.tsv"), "-output",
                os.path.join(OUT, "config1_table_scores.tsv"), "-model", "unused", "-device", "cpu", "-batchSize", "37"]
    zss.main()
    full = os.path.join(OUT, "config1_table_scores.tsv")          # keep the score column as main() spelled it, not the windows again
    pd.read_csv(full, sep="\t", dtype=str)[["chr", "pos", "ref", "alt", "zeroShotScore"]].to_csv(
        os.path.join(OUT, "config1_scores.tsv"), sep="\t", index=False)
    os.remove(full)


def main():
    os.makedirs(OUT, exist_ok=True)
    _install_stubs()
    zss = _load(os.path.join(REF, "src", "zero_shot_score.py"), "ref_zero_shot_score")
    if "--config1" in sys.argv:
        config1(zss)
        return
    zse = _load(os.path.join(REF, "src", "zero-shot-eval.py"), "ref_zero_shot_eval")
    txg = _load(os.path.join(REF, "src", "train_XGBoost.py"), "ref_train_xgboost")
    model, tok = tiny_model()
    zss.load_model_and_tokenizer = lambda model_dir, device: (model, tok)

    # ---- src/zero_shot_score.py main(): table -> TSV, table -> BED, VCF + FASTA -> per-record INFO ----------------------
    table = os.path.join(HERE, "example_snp.tsv")
    for extra, name in (([], "table_scores.tsv"), (["-outBED"], "table_scores.bed")):
        sys.argv = ["zero_shot_score.py", "-input-table", table, "-output", os.path.join(OUT, name), "-model", "unused",
                    "-device", "cpu", "-batchSize", "64"] + extra
        zss.main()
    vcf_in, fasta = os.path.join(HERE, "example_maize_snp.vcf"), os.path.join(HERE, "example_genome.fa.gz")
    args = SimpleNamespace(inputVCF=vcf_in, inputFasta=fasta, tokenIdx=255)
    sequences, record_indices = zss.seq_from_vcf(args)
    assert all(len(s) == 512 for s in sequences)
    np.savez_compressed(os.path.join(OUT, "vcf_windows.npz"),
                        windows=np.frombuffer("".join(sequences).encode(), dtype=np.uint8).reshape(len(sequences), 512),
                        record_indices=np.asarray(record_indices, dtype=np.int64))
    # window rule at the chromosome ends and for another tokenIdx, on a small genome
    small_fa = os.path.join(OUT, "small_genome.fa")
    rng = np.random.default_rng(11)
    chrom = {"c1": "".join(rng.choice(list("ACGTacgtN"), size=700)), "c2": "".join(rng.choice(list("ACGT"), size=300))}
    with open(small_fa, "w") as f:
        for k, v in chrom.items():
            f.write(f">{k} some description\n")
            f.write("\n".join(v[i:i + 60] for i in range(0, len(v), 60)) + "\n")
    small_vcf = os.path.join(OUT, "small.vcf")
    rows = [("c1", p) for p in (1, 2, 100, 255, 256, 257, 300, 444, 445, 446, 699, 700)] + [("c2", p) for p in (1, 45, 150, 299, 300)]
    with open(small_vcf, "w") as f:
        f.write("##fileformat=VCFv4.2\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        for c, p in rows:
            ref = chrom[c][p - 1].upper()
            ref = ref if ref in "ACGT" else "A"
            f.write(f"{c}\t{p}\t.\t{ref}\t{'CAGT'['ACGT'.index(ref)]}\t.\tPASS\t.\n")
    edge = {}
    for tidx in (255, 100, 0, 511):
        seqs, ridx = zss.seq_from_vcf(SimpleNamespace(inputVCF=small_vcf, inputFasta=small_fa, tokenIdx=tidx))
        edge[str(tidx)] = {"windows": seqs, "record_indices": [int(i) for i in ridx]}
    with open(os.path.join(OUT, "small_windows.json"), "w") as f:
        json.dump(edge, f)
    WRITTEN.clear()
    sys.argv = ["zero_shot_score.py", "-input-vcf", vcf_in, "-input-fasta", fasta, "-output", os.path.join(OUT, "_scratch.vcf"),
                "-model", "unused", "-device", "cpu", "-batchSize", "64"]
    zss.main()
    os.remove(os.path.join(OUT, "_scratch.vcf"))
    with open(os.path.join(OUT, "vcf_info.json"), "w") as f:
        json.dump([{"record": int(i), "plantCAD_zero_shot": v} for i, v in WRITTEN], f, indent=0)

    # ---- src/zero-shot-eval.py ------------------------------------------------------------------------------------------
    from torch.utils.data import DataLoader
    rng = np.random.default_rng(5)
    L = 96
    seqs = pd.Series(["".join(rng.choice(list("ACGT"), size=L)) for _ in range(7)])
    seqs[2] = seqs[2][:40] + "n" + seqs[2][41:]                 # an unknown base at a masked position
    seqs[4] = seqs[4].lower()
    z = {"seqs": np.array(seqs.tolist())}
    z["masked_single_40"] = zse._masked_probs(model, tok, DataLoader(zse.SingleMaskDataset(seqs, tok, 40), batch_size=3, shuffle=False),
                                              "cpu")
    z["masked_multi_40_41_42"] = zse._masked_probs(
        model, tok, DataLoader(zse.MultiMaskDataset(seqs, tok, [40, 41, 42]), batch_size=4, shuffle=False), "cpu")
    z["masked_multi_unsorted_70_5_41"] = zse._masked_probs(
        model, tok, DataLoader(zse.MultiMaskDataset(seqs, tok, [70, 5, 41]), batch_size=4, shuffle=False), "cpu")
    z["unmasked"] = zse._unmasked_probs(seqs, tok, model, "cpu", 3)
    for pos_name, positions in (("40", [40]), ("40_41_42", [40, 41, 42])):
        z[f"true_tokens_{pos_name}"] = zse._compute_true_tokens_from_seq(seqs, positions)
    # metrics on seeded random probabilities (no model involved)
    n = 60
    mseqs = pd.Series(["".join(rng.choice(list("ACGTacgtN"), size=12)) for _ in range(n)])
    labels = rng.integers(0, 2, size=n)
    df = pd.DataFrame({"sequence": mseqs, "label": labels})
    def noisy_truth(tokens):          # logits that favour the true base at ~70 % of the positions: accuracies land mid-range
        lg = rng.normal(size=(len(tokens), 4)) * 2
        for k, t in enumerate(tokens):
            if t in "ACGT" and rng.random() < 0.7:
                lg[k, "ACGT".index(t)] += 6
        return torch.softmax(torch.from_numpy(lg), dim=1).numpy().astype(np.float32)

    p1 = noisy_truth([s[5].upper() for s in mseqs])
    p3 = noisy_truth([s[i].upper() for s in mseqs for i in (4, 5, 6)])
    z["m_seqs"], z["m_labels"], z["m_probs1"], z["m_probs3"] = np.array(mseqs.tolist()), labels, p1, p3
    z["m_auroc_idx5"] = zse._compute_auroc(df, p1, 5, "sequence")
    z["m_refprob_idx5"] = zse._refprob_scores(df, p1, 5, "sequence")
    tt1 = zse._compute_true_tokens_from_seq(mseqs, [5])
    tt3 = zse._compute_true_tokens_from_seq(mseqs, [4, 5, 6])
    z["m_token_acc1"] = zse._metric_token_accuracy(p1, tt1)
    z["m_token_acc3"] = zse._metric_token_accuracy(p3, tt3)
    z["m_motif_acc3"] = zse._metric_motif_accuracy(p3, tt3, 3)
    z["m_avg_trueprob3"] = zse._avg_trueprob_scores(p3, tt3, 3)
    # SV boundary score
    nsv, Lsv, fl = 9, 64, 5
    ref_seqs = ["".join(rng.choice(list("ACGT"), size=Lsv)) for _ in range(nsv)]
    mut_seqs = ["".join(rng.choice(list("ACGTacgtN"), size=Lsv)) for _ in range(nsv)]
    left = rng.integers(fl + 1, Lsv // 2, size=nsv)
    right = rng.integers(Lsv // 2, Lsv - fl, size=nsv)
    sv = pd.DataFrame({"RefSeq": ref_seqs, "MutSeq": mut_seqs, "left": left, "right": right, "label": rng.integers(0, 2, size=nsv)})
    rp = torch.softmax(torch.from_numpy(rng.normal(size=(nsv, Lsv, 4)) * 3), dim=-1).numpy().astype(np.float32)
    mp = torch.softmax(torch.from_numpy(rng.normal(size=(nsv, Lsv, 4)) * 3), dim=-1).numpy().astype(np.float32)
    rp[0, :, 0] = 0.0                                        # exercises the 1e-12 floor
    z["sv_mut_seqs"], z["sv_left"], z["sv_right"], z["sv_ref_probs"], z["sv_mut_probs"] = np.array(mut_seqs), left, right, rp, mp
    z["sv_scores_fl5"] = zse._sv_llr_boundary(sv, rp, mp, fl)
    z["sv_scores_fl2"] = zse._sv_llr_boundary(sv, rp, mp, 2)
    np.savez_compressed(os.path.join(OUT, "zero_shot_eval.npz"), **z)

    # ---- src/zero-shot-eval.py: the four ZeroShotEval commands end to end on a local dataset ---------------------------
    import contextlib
    import io
    n_ds, L_ds = 24, 96
    ds_seqs = ["".join(rng.choice(list("ACGT"), size=L_ds)) for _ in range(n_ds)]
    ds_seqs[3] = ds_seqs[3][:40] + "N" + ds_seqs[3][41:]
    ds_seqs[7] = ds_seqs[7].lower()
    ds_df = pd.DataFrame({"sequence": ds_seqs, "label": rng.integers(0, 2, size=n_ds)})
    sv_df = pd.DataFrame({"RefSeq": ["".join(rng.choice(list("ACGT"), size=L_ds)) for _ in range(n_ds)],
                          "MutSeq": ["".join(rng.choice(list("ACGTN"), size=L_ds, p=[0.24, 0.24, 0.24, 0.24, 0.04])) for _ in range(n_ds)],
                          "left": rng.integers(6, L_ds // 2, size=n_ds), "right": rng.integers(L_ds // 2, L_ds - 6, size=n_ds),
                          "label": rng.integers(0, 2, size=n_ds), "Left5_Positions": ["x"] * n_ds})
    ds_df.to_csv(os.path.join(OUT, "eval_dataset.tsv"), sep="\t", index=False)
    sv_df.to_csv(os.path.join(OUT, "eval_sv_dataset.tsv"), sep="\t", index=False)

    class Split:
        def __init__(self, df):
            self.df = df

        def to_pandas(self):
            return self.df.copy()

    zse.load_dataset = lambda repo_id, task: {"valid": Split(sv_df if task == "sv" else ds_df)}
    zse._require_cuda = lambda device: device
    zse._load_model = lambda model_name, device: (model, tok)
    runs = {}

    def run(name, fn, **kw):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            fn(repo_id="local", task="sv" if name == "sv_effect" else "t", device="cpu", **kw)
        runs[name] = {"stdout": buf.getvalue()}

    ev = zse.ZeroShotEval()
    run("evo_cons", ev.evo_cons, token_idx=40, batch_size=5, save_logits=os.path.join(OUT, "eval_evo_cons_logits.tsv"),
        metrics_json=os.path.join(OUT, "eval_evo_cons_metrics.json"))
    run("motif_acc", ev.motif_acc, mask_idx=(40, 41, 42), motif_len=3, batch_size=5,
        save_logits=os.path.join(OUT, "eval_motif_logits.tsv"), metrics_json=os.path.join(OUT, "eval_motif_metrics.json"))
    run("core_noncore", ev.core_noncore, mask_idx=(40, 41, 42), motif_len=3, batch_size=7,
        metrics_json=os.path.join(OUT, "eval_core_noncore_metrics.json"))
    run("sv_effect", ev.sv_effect, batch_size=5, flanking=5, output=os.path.join(OUT, "eval_sv_scored.tsv"))
    with open(os.path.join(OUT, "eval_runs.json"), "w") as f:
        json.dump(runs, f, indent=1)

    # ---- src/train_XGBoost.py -------------------------------------------------------------------------------------------
    loader = txg.create_dataloader(seqs.tolist(), tok, 3)
    emb = txg.extract_embeddings(model, loader, "cpu", 40)
    np.savez_compressed(os.path.join(OUT, "embeddings.npz"), seqs=np.array(seqs.tolist()), token_idx=40, averaged=emb)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
