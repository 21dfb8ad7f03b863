"""Drop-in for the reference's ``src/zero-shot-eval.py`` (PlantCAD2 zero-shot evaluation) on the B200 engine
(SURVEY.md 8f item 3): the model-facing helpers -- masked-position probabilities for single / multi-mask inputs
(`_masked_probs`, :129-140), per-position probabilities for whole windows (`_unmasked_probs`, :143-178), the
structural-variant boundary score (`_sv_llr_boundary`, :181-243) -- the metric helpers (:246-316) and the four commands of
its ``ZeroShotEval`` class (:319-530: ``evo_cons``, ``motif_acc``, ``sv_effect``, ``core_noncore``) with the reference's
argument names, printed lines and output files:

    python -m plantcaduceus_b200.zero_shot_eval evo_cons --repo_id <dataset> --task <task> --model <dir|preset> --token_idx 255

Differences in how the work is done: sequences travel as ASCII bytes and are tokenised on the device; for masked
inputs the LM head runs only at the masked positions (``pcad_score_masked``) instead of producing ``[B, L, V]``
logits and selecting afterwards; ``sv_effect`` keeps the two ``[N, L, 4]`` probability tensors on the device and gathers
the 2 * flanking values per row there; the scores and metrics are vectorised numpy instead of per-row Python loops.
``repo_id`` may be a Hugging Face dataset name (needs the hub or a local cache) or a local file / directory
(``_load_split``).  Every function is checked against the outputs of the reference's own code
(tests/golden/reference_run/, tests/test_reference_run.py).
"""
from __future__ import annotations

import ast
import inspect
import json
import logging
import os
import sys
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import genome_io as gio

logger = logging.getLogger(__name__)

NUCLEOTIDES = ("A", "C", "G", "T")


def _ascii_matrix(tokenizer, sequences: Sequence[str]) -> np.ndarray:
    seqs = [str(s) for s in sequences]
    if not seqs:
        return np.zeros((0, 0), dtype=np.uint8)
    L = len(seqs[0])
    for s in seqs:
        if len(s) != L:
            raise ValueError(f"All sequences must have same length; got {len(s)} vs {L}")
    return tokenizer.windows_to_ascii(seqs, L)


def masked_probs(model, tokenizer, sequences: Sequence[str], mask_idx: Sequence[int], batch_size: int = 64) -> np.ndarray:
    """Softmax over a,c,g,t at every masked position: float32 ``[N * len(mask_idx), 4]``, sequence-major then in
    increasing position order -- the order ``torch.masked_select`` yields in the reference (:136-137).
    ``mask_idx`` may be one index (SingleMaskDataset) or several (MultiMaskDataset)."""
    mask_idx = sorted(int(i) for i in ([mask_idx] if np.isscalar(mask_idx) else mask_idx))
    ascii_mat = _ascii_matrix(tokenizer, sequences)
    n, L = ascii_mat.shape
    if n and mask_idx and mask_idx[-1] >= L:
        raise AssertionError("mask index out of range")
    out = np.zeros((n * len(mask_idx), 4), dtype=np.float32)
    pos_row = np.asarray(mask_idx, dtype=np.int32)
    for s in range(0, n, batch_size):
        ids = tokenizer.encode_bytes(ascii_mat[s:s + batch_size])
        ids[:, pos_row] = tokenizer.mask_token_id
        pos = np.repeat(pos_row[None, :], len(ids), axis=0)
        logits4 = model.score_masked(torch.from_numpy(ids), torch.from_numpy(pos)).cpu().numpy()
        out[s * len(mask_idx):(s + len(ids)) * len(mask_idx)] = gio.softmax4(logits4.reshape(-1, 4))
    return out


def unmasked_probs(model, tokenizer, sequences: Sequence[str], batch_size: int = 16, on_device: bool = False):
    """Per-position probabilities over A,C,G,T for each (unmasked) sequence: float32 ``[N, L, 4]`` (:143-178).
    ``on_device=True`` returns a torch tensor that stays on the model's device (at L = 8192 the ``[N, L, 4]`` array is what
    crosses PCIe otherwise; ``sv_llr_boundary`` accepts it and gathers its 2 * flanking values per row there)."""
    ascii_mat = _ascii_matrix(tokenizer, sequences)
    n, L = ascii_mat.shape
    v = tokenizer.get_vocab()
    idxs = [v[c] for c in "acgt"]
    if on_device:
        out_dev = torch.zeros((n, L, 4), dtype=torch.float32, device=model.device)
        sel = torch.tensor(idxs, device=model.device)
        for s in range(0, n, batch_size):
            ids = torch.from_numpy(tokenizer.encode_bytes(ascii_mat[s:s + batch_size]).astype(np.int64))
            logits = model(input_ids=ids.to(model.device)).logits
            out_dev[s:s + len(ids)] = torch.softmax(logits.index_select(-1, sel).float(), dim=-1)
        return out_dev
    out = np.zeros((n, L, 4), dtype=np.float32)
    for s in range(0, n, batch_size):
        ids = torch.from_numpy(tokenizer.encode_bytes(ascii_mat[s:s + batch_size]).astype(np.int64))
        logits = model(input_ids=ids.to(model.device)).logits.cpu().numpy()[..., idxs]
        out[s:s + len(ids)] = gio.softmax4(logits.reshape(-1, 4)).reshape(len(ids), L, 4)
    return out


def sv_llr_boundary(left: Sequence[int], right: Sequence[int], mut_seqs: Sequence[str], ref_probs: np.ndarray,
                    mut_probs: np.ndarray, flanking: int) -> np.ndarray:
    """-mean log(p_mut / p_ref) over the 2*flanking boundary positions of each structural variant (:181-243).
    ``left`` / ``right`` are the 1-based breakpoints in the reference window; the mutated window is centred at
    ``L // 2``.  Reference positions: ``[left-flanking .. left-1]`` and ``[right+1 .. right+flanking]`` (1-based);
    mutated positions: ``[c-flanking .. c-1]`` and ``[c .. c+flanking-1]`` (0-based).  The channel is the mutated base
    at the mutated position; non-ACGT bases contribute 0; probabilities are floored at 1e-12."""
    n = len(mut_seqs)
    L = ref_probs.shape[1]
    c = L // 2
    if isinstance(ref_probs, torch.Tensor) or isinstance(mut_probs, torch.Tensor):
        return _sv_llr_boundary_device(left, right, mut_seqs, ref_probs, mut_probs, flanking)
    k = np.arange(flanking)
    left = np.asarray(left, dtype=np.int64)
    right = np.asarray(right, dtype=np.int64)
    # 0-based reference positions, [n, 2*flanking] in (left block, right block) order
    ref_pos = np.concatenate([(left[:, None] - 1) - (flanking - 1) + k[None, :] - 1, (right[:, None] + 1) + k[None, :] - 1], axis=1)
    mut_pos = np.concatenate([c - flanking + k, c + k])[None, :].repeat(n, axis=0)
    centre = np.array([list(str(s)[c - flanking:c + flanking].upper().ljust(2 * flanking, "N")) for s in mut_seqs])
    chan = np.full(centre.shape, -1, dtype=np.int64)
    for j, b in enumerate(NUCLEOTIDES):
        chan[centre == b] = j
    valid = chan >= 0
    rows = np.arange(n)[:, None].repeat(2 * flanking, axis=1)
    ch = np.where(valid, chan, 0)
    r = ref_probs[rows, ref_pos, ch]
    m = mut_probs[rows, mut_pos, ch]
    vals = np.where(valid, np.log(np.maximum(m, 1e-12) / np.maximum(r, 1e-12)), 0.0)
    return -vals.mean(axis=1)


def _sv_llr_boundary_device(left, right, mut_seqs, ref_probs, mut_probs, flanking: int) -> np.ndarray:
    """``sv_llr_boundary`` with the probability tensors resident on the device: the 2 * flanking (position, channel) pairs of
    every row are gathered there and only the ``[n]`` scores come back."""
    dev = ref_probs.device if isinstance(ref_probs, torch.Tensor) else mut_probs.device
    ref_probs = torch.as_tensor(ref_probs, device=dev)
    mut_probs = torch.as_tensor(mut_probs, device=dev)
    n, L = len(mut_seqs), ref_probs.shape[1]
    c = L // 2
    k = torch.arange(flanking, device=dev)
    left_t = torch.as_tensor(np.array(left, dtype=np.int64), device=dev)
    right_t = torch.as_tensor(np.array(right, dtype=np.int64), device=dev)
    ref_pos = torch.cat([(left_t[:, None] - 1) - (flanking - 1) + k[None, :] - 1, (right_t[:, None] + 1) + k[None, :] - 1], dim=1)
    mut_pos = torch.cat([c - flanking + k, c + k])[None, :].expand(n, -1)
    centre = np.array([list(str(s)[c - flanking:c + flanking].upper().ljust(2 * flanking, "N")) for s in mut_seqs])
    chan = np.full(centre.shape, -1, dtype=np.int64)
    for j, b in enumerate(NUCLEOTIDES):
        chan[centre == b] = j
    chan_t = torch.as_tensor(chan, device=dev)
    valid = chan_t >= 0
    ch = torch.where(valid, chan_t, torch.zeros_like(chan_t))
    rows = torch.arange(n, device=dev)[:, None].expand(-1, 2 * flanking)
    r = ref_probs[rows, ref_pos, ch]
    m = mut_probs[rows, mut_pos, ch]
    vals = torch.where(valid, torch.log(torch.clamp(m, min=1e-12) / torch.clamp(r, min=1e-12)), torch.zeros_like(m))
    return (-vals.mean(dim=1)).cpu().numpy()


def compute_true_tokens_from_seq(sequences: Sequence[str], positions: List[int]) -> np.ndarray:
    """Upper-cased base of every sequence at every position, sequence-major (:246-251)."""
    return np.array([str(seq)[i].upper() for seq in sequences for i in positions])


# ---------------------------------------------------------------------------------------------------------------------
# metrics (reference :254-316), vectorised
# ---------------------------------------------------------------------------------------------------------------------
def _base_index(tokens) -> np.ndarray:
    """Index into A,C,G,T of every (upper-case) token, -1 for anything else."""
    tokens = np.asarray(tokens)
    idx = np.full(tokens.shape, -1, dtype=np.int64)
    for j, b in enumerate(NUCLEOTIDES):
        idx[tokens == b] = j
    return idx


def refprob_scores(df, probs: np.ndarray, token_idx: int, seq_col: str) -> np.ndarray:
    """Probability the model gives the REFERENCE base at the masked index, one per row; 0 where that base is not A/C/G/T
    (:274-282)."""
    ref = _base_index(df[seq_col].str[token_idx].str.upper().to_numpy(dtype=str))
    p = np.asarray(probs).reshape(len(df), -1)
    scores = np.zeros(len(df), dtype=float)
    valid = ref >= 0
    scores[valid] = p[np.flatnonzero(valid), ref[valid]]
    return scores


def compute_auroc(df, probs: np.ndarray, token_idx: int, seq_col: str) -> float:
    """AUROC of ``label`` against ``refprob_scores`` (:260-272)."""
    from sklearn.metrics import auc, roc_curve
    fpr, tpr, _ = roc_curve(df["label"].astype(int), refprob_scores(df, probs, token_idx, seq_col))
    return float(auc(fpr, tpr))


def metric_token_accuracy(probs: np.ndarray, true_tokens: np.ndarray) -> float:
    """Fraction of positions with a known true base whose argmax is that base (:254-261); 0.0 if none is known."""
    true = _base_index(true_tokens)
    valid = true >= 0
    if not valid.any():
        return 0.0
    return float((np.asarray(probs).argmax(axis=1)[valid] == true[valid]).mean())


def metric_motif_accuracy(probs: np.ndarray, true_tokens: np.ndarray, motif_len: int) -> float:
    """Fraction of fully known motifs (groups of ``motif_len`` consecutive rows) predicted exactly (:264-274)."""
    true = _base_index(true_tokens)
    assert len(true) % motif_len == 0, "total masked positions not divisible by motif_len"
    pred = np.asarray(probs).argmax(axis=1).reshape(-1, motif_len)
    true = true.reshape(-1, motif_len)
    valid = (true >= 0).all(axis=1)
    if not valid.any():
        return 0.0
    return float((pred[valid] == true[valid]).all(axis=1).mean())


def avg_trueprob_scores(probs: np.ndarray, true_tokens: np.ndarray, motif_len: int) -> np.ndarray:
    """Mean probability of the true base over each example's ``motif_len`` masked positions, unknown bases counting 0
    (:285-305)."""
    true = _base_index(true_tokens)
    assert len(true) % motif_len == 0, "total masked positions not divisible by motif_len"
    probs = np.asarray(probs)
    token_probs = np.zeros(probs.shape[0], dtype=float)
    valid = true >= 0
    token_probs[valid] = probs[np.flatnonzero(valid), true[valid]]
    return token_probs.reshape(-1, motif_len).mean(axis=1)


# the reference's private spellings, for callers that import them
_compute_true_tokens_from_seq = compute_true_tokens_from_seq
_refprob_scores, _compute_auroc = refprob_scores, compute_auroc
_metric_token_accuracy, _metric_motif_accuracy, _avg_trueprob_scores = metric_token_accuracy, metric_motif_accuracy, avg_trueprob_scores


# ---------------------------------------------------------------------------------------------------------------------
# the commands (reference class ZeroShotEval, :319-530)
# ---------------------------------------------------------------------------------------------------------------------
def _require_cuda(device: str) -> str:
    """The reference refuses to run on a CPU (:29-41); so does the engine (it has no CPU path)."""
    if not (torch.cuda.is_available() and str(device).startswith("cuda")):
        raise RuntimeError("CUDA is required for zero-shot evaluation; CPU is not supported. "
                           "Set -device cuda:0 and ensure a CUDA GPU is available.")
    return device


def _load_model(model_name: str, device: str):
    """``AutoModelForMaskedLM.from_pretrained(model_name, torch_dtype=bf16)`` + tokenizer, on ``device`` (:54-72).  A local
    checkpoint directory is loaded; any other name must be a preset and gets random-initialised weights (no hub offline)."""
    from .modeling import CaduceusForMaskedLM
    from .tokenizer import CharDNATokenizer
    if os.path.isdir(str(model_name)):
        model = CaduceusForMaskedLM.from_pretrained(model_name, torch_dtype=torch.bfloat16)
        tok = CharDNATokenizer.from_pretrained(model_name)
    else:
        name = str(model_name).split("/")[-1]
        logger.warning("%r is not a local directory: random-initialised weights of the preset %r", model_name, name)
        model = CaduceusForMaskedLM.from_random(name, seed=0, torch_dtype=torch.bfloat16)
        tok = CharDNATokenizer()
    model.set_tokenizer(tok)
    model.to(device)
    return model, tok


def _read_frame(path: str):
    import pandas as pd
    low = path.lower()
    if low.endswith(".parquet"):
        return pd.read_parquet(path)
    if low.endswith((".csv", ".csv.gz")):
        return pd.read_csv(path)
    return pd.read_csv(path, sep="\t")


def _load_split(repo_id: str, task: str, split: str):
    """``load_dataset(repo_id, task)[split].to_pandas()`` (:341-342).  ``repo_id`` may also be a local file (parquet / TSV /
    CSV) or a directory holding ``<task>/<split>.*``, ``<task>/<split>-*.parquet`` or ``<task>_<split>.*``."""
    import glob
    if os.path.isfile(repo_id):
        return _read_frame(repo_id)
    if os.path.isdir(repo_id):
        pats = [os.path.join(repo_id, task, f"{split}.*"), os.path.join(repo_id, task, f"{split}-*.parquet"),
                os.path.join(repo_id, f"{task}_{split}.*"), os.path.join(repo_id, task, "**", f"{split}-*.parquet")]
        for pat in pats:
            hits = sorted(glob.glob(pat, recursive=True))
            if hits:
                import pandas as pd
                return pd.concat([_read_frame(h) for h in hits], ignore_index=True) if len(hits) > 1 else _read_frame(hits[0])
        raise FileNotFoundError(f"no file for task {task!r} split {split!r} under {repo_id}")
    from datasets import load_dataset
    return load_dataset(repo_id, task)[split].to_pandas()


class ZeroShotEval:
    """The reference's four commands, same names and arguments."""

    @staticmethod
    def _probs_masked(df, seq_column, positions, model, device, batch_size, save_logits, logits_path):
        import pandas as pd
        if logits_path is not None:
            return pd.read_csv(logits_path, sep="\t").values
        dev = _require_cuda(device)
        model_, tok = _load_model(model, dev)
        probs = masked_probs(model_, tok, df[seq_column].tolist(), positions, batch_size)
        if save_logits:
            pd.DataFrame(probs, columns=list(NUCLEOTIDES)).to_csv(save_logits, sep="\t", index=False)
            logger.info(f"Saved logits TSV to {save_logits}")
        return probs

    def evo_cons(self, repo_id: str, task: str, split: str = "valid", model: str = "kuleshov-group/PlantCAD2-Small-l24-d0768",
                 device: str = "cuda:0", token_idx: int = 255, batch_size: int = 128, seq_column: str = "sequence",
                 save_logits: Optional[str] = None, logits_path: Optional[str] = None, metrics_json: Optional[str] = None) -> None:
        """Masked-token probabilities at one index; AUROC / AUPRC of ``label`` against the reference base's probability."""
        from sklearn.metrics import average_precision_score
        logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")
        df = _load_split(repo_id, task, split)
        probs = self._probs_masked(df, seq_column, [int(token_idx)], model, device, batch_size, save_logits, logits_path)
        assert probs.shape[0] == len(df), f"Row mismatch: probs={probs.shape[0]} examples={len(df)}"
        roc_auc = compute_auroc(df, probs, token_idx, seq_column)
        auprc = float(average_precision_score(df["label"].astype(int).to_numpy(), refprob_scores(df, probs, token_idx, seq_column)))
        print(f"AUROC\t{roc_auc:.6f}")
        print(f"AUPRC\t{auprc:.6f}")
        if metrics_json:
            with open(metrics_json, "w") as f:
                json.dump({"auroc": roc_auc, "auprc": auprc, "token_idx": token_idx}, f, indent=2)

    def motif_acc(self, repo_id: str, task: str, split: str = "valid", model: str = "kuleshov-group/PlantCAD2-Small-l24-d0768",
                  device: str = "cuda:0", mask_idx: Sequence[int] = (255, 256, 257), motif_len: int = 3, batch_size: int = 128,
                  seq_column: str = "sequence", save_logits: Optional[str] = None, logits_path: Optional[str] = None,
                  metrics_json: Optional[str] = None) -> None:
        """Multi-position masked probabilities; token and motif accuracy."""
        logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")
        df = _load_split(repo_id, task, split)
        positions = [int(x) for x in mask_idx]
        assert len(positions) == motif_len, "mask_idx count must equal motif_len"
        probs = self._probs_masked(df, seq_column, positions, model, device, batch_size, save_logits, logits_path)
        expected = len(df) * len(positions)
        assert probs.shape[0] == expected, f"Row mismatch: probs={probs.shape[0]} expected={expected}"
        true_tokens = compute_true_tokens_from_seq(df[seq_column], positions)
        token_acc = metric_token_accuracy(probs, true_tokens)
        motif = metric_motif_accuracy(probs, true_tokens, motif_len)
        print(f"token_accuracy\t{token_acc:.6f}")
        print(f"motif_accuracy\t{motif:.6f}")
        if metrics_json:
            with open(metrics_json, "w") as f:
                json.dump({"token_accuracy": token_acc, "motif_accuracy": motif}, f, indent=2)

    def sv_effect(self, repo_id: str, task: str, split: str = "valid", model: str = "kuleshov-group/PlantCAD2-Small-l24-d0768",
                  device: str = "cuda:0", batch_size: int = 64, flanking: int = 5, output: Optional[str] = None,
                  save_ref_logits: Optional[str] = None, save_mut_logits: Optional[str] = None) -> None:
        """Structural-variant effect: unmasked probabilities of RefSeq and MutSeq, boundary log-ratio score, AUPRC."""
        from sklearn.metrics import average_precision_score
        logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")
        df = _load_split(repo_id, task, split)
        missing = [c for c in ["RefSeq", "MutSeq", "left", "right", "label"] if c not in df.columns]
        if missing:
            raise KeyError(f"Missing required columns: {missing}")
        dev = _require_cuda(device)
        model_, tok = _load_model(model, dev)
        ref_probs = unmasked_probs(model_, tok, df["RefSeq"].astype(str).tolist(), batch_size, on_device=True)
        mut_probs = unmasked_probs(model_, tok, df["MutSeq"].astype(str).tolist(), batch_size, on_device=True)
        if save_ref_logits:
            np.savez_compressed(save_ref_logits, logits=ref_probs.cpu().numpy())
        if save_mut_logits:
            np.savez_compressed(save_mut_logits, logits=mut_probs.cpu().numpy())
        scores = sv_llr_boundary(df["left"].to_numpy(), df["right"].to_numpy(), df["MutSeq"].tolist(), ref_probs, mut_probs, flanking)
        auprc = float(average_precision_score(df["label"].astype(int).to_numpy(), scores))
        print(f"AUPRC\t{auprc:.6f}")
        if output:
            out_df = df.copy()
            out_df["score"] = scores
            out_df = out_df.drop(columns=["Left5_Positions", "Right5_Positions"], errors="ignore")
            out_df.to_csv(output, sep="\t", index=False)

    def core_noncore(self, repo_id: str, task: str, split: str = "valid", model: str = "kuleshov-group/PlantCAD2-Small-l24-d0768",
                     device: str = "cuda:0", mask_idx: Sequence[int] = (255, 256, 257), motif_len: int = 3, batch_size: int = 128,
                     seq_column: str = "sequence", label_column: str = "label", save_logits: Optional[str] = None,
                     logits_path: Optional[str] = None, metrics_json: Optional[str] = None) -> None:
        """Core vs non-core classification by the mean true-base probability over the masked positions; AUROC / AUPRC."""
        from sklearn.metrics import auc, average_precision_score, roc_curve
        logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")
        df = _load_split(repo_id, task, split)
        positions = [int(x) for x in mask_idx]
        assert len(positions) == motif_len, "mask_idx count must equal motif_len"
        probs = self._probs_masked(df, seq_column, positions, model, device, batch_size, save_logits, logits_path)
        expected = len(df) * len(positions)
        assert probs.shape[0] == expected, f"Row mismatch: probs={probs.shape[0]} expected={expected}"
        scores = avg_trueprob_scores(probs, compute_true_tokens_from_seq(df[seq_column], positions), motif_len)
        y_true = df[label_column].astype(int).to_numpy()
        fpr, tpr, _ = roc_curve(y_true, scores)
        roc_auc = float(auc(fpr, tpr))
        auprc = float(average_precision_score(y_true, scores))
        print(f"AUROC\t{roc_auc:.6f}")
        print(f"AUPRC\t{auprc:.6f}")
        if metrics_json:
            with open(metrics_json, "w") as f:
                json.dump({"auroc": roc_auc, "auprc": auprc}, f, indent=2)


def _parse_value(text: str):
    """python-fire's reading of a flag value: a Python literal when it parses as one, else the string."""
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError):
        return text


def main(argv: Optional[Sequence[str]] = None) -> int:
    """``fire.Fire(ZeroShotEval)`` without the dependency: ``<command> [positional ...] [--name value | --name=value]``;
    flag names take ``-`` or ``_`` (``--mask-idx=255,256,257``), values are Python literals where they parse as such."""
    argv = list(sys.argv[1:] if argv is None else argv)
    commands = [m for m in ("evo_cons", "motif_acc", "sv_effect", "core_noncore")]
    if not argv or argv[0].replace("-", "_") not in commands:
        print("usage: python -m plantcaduceus_b200.zero_shot_eval {" + ",".join(commands) + "} --repo_id R --task T [--flag value ...]",
              file=sys.stderr)
        return 2
    fn = getattr(ZeroShotEval(), argv[0].replace("-", "_"))
    params = list(inspect.signature(fn).parameters)
    kwargs, positional = {}, []
    i = 1
    while i < len(argv):
        a = argv[i]
        if a.startswith("--"):
            name, eq, val = a[2:].partition("=")
            name = name.replace("-", "_")
            if name not in params:
                print(f"unknown flag --{name} for {argv[0]}", file=sys.stderr)
                return 2
            if not eq:
                if i + 1 < len(argv) and not argv[i + 1].startswith("--"):
                    val, i = argv[i + 1], i + 1
                else:
                    val = "True"
            kwargs[name] = _parse_value(val)
        else:
            positional.append(_parse_value(a))
        i += 1
    for name, val in zip([p for p in params if p not in kwargs], positional):
        kwargs[name] = val
    for name in ("repo_id", "task", "model", "split", "seq_column", "label_column", "device"):
        if name in kwargs:
            kwargs[name] = str(kwargs[name])
    fn(**kwargs)
    return 0


if __name__ == "__main__":
    sys.exit(main())
