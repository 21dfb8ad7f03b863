"""Zero-shot SNP scoring on the B200 engine -- drop-in for the reference's ``src/zero_shot_score.py``.

Same command line (single-dash long flags, reference :14-37), same inputs (SNP table with
``ref, alt, sequences`` columns, or VCF + FASTA), same outputs (TSV with a ``zeroShotScore`` column, headerless
BED with ``start = pos-1``, or VCF with ``INFO/plantCAD_zero_shot``), same function names for the steps.
What changes is how the steps run:

  reference                                              here
  ---------                                              ----
  per-sequence HF tokenizer call in the main process     windows travel as ASCII bytes; tokenise + mask on the GPU
  model(input_ids).logits for all 512 x 8 positions      LM head only at the masked position, 4 columns
  Biopython / PyVCF3 / per-row Python loops              byte-level FASTA/VCF readers, vectorised numpy scoring
  one process, one GPU                                   under torchrun: windows sharded over the ranks (one per GPU),
                                                         per-variant scores gathered over NCCL, rank 0 writes

    python -m plantcaduceus_b200.zero_shot_score -input-table examples/example_snp.tsv -output out.tsv -model <dir|preset>
"""
from __future__ import annotations

import argparse
import logging
import sys
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import genome_io as gio
from .tokenizer import CharDNATokenizer


def parse_args(argv: Optional[Sequence[str]] = None):
    """Flag spellings and defaults of the reference (src/zero_shot_score.py:14-37) plus -seed / -dtype for
    random-initialised presets (no checkpoints offline)."""
    p = argparse.ArgumentParser(description="PlantCaduceus zero-shot variant scoring (B200 engine)")
    g = p.add_mutually_exclusive_group(required=True)
    g.add_argument("-input-table", dest="inputDF", type=str, default=None,
                   help="tab-separated file with columns ref, alt, sequences")
    g.add_argument("-input-vcf", dest="inputVCF", type=str, default=None, help="input VCF")
    p.add_argument("-input-fasta", dest="inputFasta", type=str, default=None, help="genome FASTA (required with -input-vcf)")
    p.add_argument("-output", dest="output", default=None, help="output path")
    p.add_argument("-outBED", action="store_true", dest="outBED", default=False,
                   help="BED output instead of TSV (only with -input-table)")
    p.add_argument("-model", dest="model", default=None,
                   help="checkpoint directory, or a preset name (PlantCaduceus_l20..l32) for random-init weights")
    p.add_argument("-device", dest="device", default="cuda:0")
    p.add_argument("-batchSize", dest="batchSize", default=128, type=int)
    p.add_argument("-numWorkers", dest="numWorkers", default=4, type=int,
                   help="accepted for compatibility; unused (as in the reference, :29 vs :104)")
    p.add_argument("-tokenIdx", dest="tokenIdx", default=255, type=int, help="index of the nucleotide to mask")
    p.add_argument("-dtype", dest="dtype", default="bfloat16", choices=["bfloat16", "float32"])
    p.add_argument("-seed", dest="seed", default=0, type=int, help="seed for random-init presets")
    args = p.parse_args(argv)
    if args.inputVCF is not None and args.inputFasta is None:
        sys.exit("-input-fasta is required with -input-vcf")
    return args


class SequenceDataset:
    """Windows as a uint8 ASCII matrix.  ``__getitem__`` keeps the reference's record layout
    (``{'sequence', 'input_ids'}`` with position tokenIdx masked, :49-62) for callers that index it; the
    scorer itself consumes whole ``ascii_batch`` slices and tokenises on the device."""

    def __init__(self, sequences: Sequence[str], tokenizer: CharDNATokenizer, tokenIdx: int):
        self.sequences = list(sequences)
        self.tokenizer = tokenizer
        self.tokenIdx = tokenIdx

    def __len__(self):
        return len(self.sequences)

    def __getitem__(self, idx):
        seq = self.sequences[idx]
        ids = self.tokenizer.encode_plus(seq, return_tensors="pt", return_attention_mask=False,
                                         return_token_type_ids=False)["input_ids"]
        ids[0, self.tokenIdx] = self.tokenizer.mask_token_id
        return {"sequence": seq, "input_ids": ids}

    def ascii_batches(self, batch_size: int):
        """Yields (start, uint8 [b, L]) over runs of equal-length windows, in input order."""
        i, n = 0, len(self.sequences)
        while i < n:
            L = len(self.sequences[i])
            j = i
            while j < n and j - i < batch_size and len(self.sequences[j]) == L:
                j += 1
            yield i, self.tokenizer.windows_to_ascii(self.sequences[i:j], L)
            i = j


def load_model_and_tokenizer(model_dir: str, device: str, dtype: str = "bfloat16", seed: int = 0):
    """``AutoModelForMaskedLM.from_pretrained(...).to(device)`` + ``AutoTokenizer.from_pretrained`` (reference :65-98).
    bf16 by default (every B200 is >= sm_80, :77-79); float32 on request for parity runs."""
    import os

    import torch

    from .modeling import CaduceusForMaskedLM
    tdtype = torch.bfloat16 if dtype == "bfloat16" else torch.float32
    logging.info(f"Loading model and tokenizer from {model_dir}")
    if model_dir is not None and os.path.isdir(model_dir):
        model = CaduceusForMaskedLM.from_pretrained(model_dir, torch_dtype=tdtype)
        tokenizer = CharDNATokenizer.from_pretrained(model_dir)
    else:
        logging.info(f"{model_dir!r} is not a directory: using random-initialised weights of that preset (seed {seed})")
        model = CaduceusForMaskedLM.from_random(model_dir or "PlantCaduceus_l32", seed=seed, torch_dtype=tdtype)
        tokenizer = CharDNATokenizer()
    model.set_tokenizer(tokenizer)
    model.to(device)
    return model, tokenizer


def create_dataloader(sequences, tokenizer, batch_size, tokenIdx):
    logging.info(f"Creating DataLoader with batch size {batch_size}")
    return SequenceDataset(sequences, tokenizer, tokenIdx), batch_size


def extract_logits(model, dataloader, device, tokenIdx, tokenizer) -> np.ndarray:
    """softmax over the a,c,g,t logits at the masked index for every window -> float32 [n, 4]
    (reference extract_logits, :107-121)."""
    import torch
    dataset, batch_size = dataloader
    logging.info("Extracting logits")
    out = np.zeros((len(dataset), 4), dtype=np.float32)
    for start, ascii_batch in dataset.ascii_batches(batch_size):
        if tokenIdx >= ascii_batch.shape[1] or tokenIdx < -ascii_batch.shape[1]:
            raise IndexError(f"tokenIdx {tokenIdx} is out of bounds for a window of length {ascii_batch.shape[1]}")
        pinned = torch.from_numpy(ascii_batch).pin_memory()
        logits4 = model.score_windows_host(pinned, tokenIdx % ascii_batch.shape[1]).numpy()
        out[start:start + len(ascii_batch)] = gio.softmax4(logits4)
    return out


def _allele_index(values) -> np.ndarray:
    lut = {n: i for i, n in enumerate(gio.NUCLEOTIDES)}
    return np.array([lut[v] for v in values], dtype=np.int64)


def zero_shot_score(snpDF, logits) -> List[float]:
    """log(p_alt / p_ref) per row of the SNP table (reference :124-134)."""
    logging.info("Calculating zero-shot scores")
    return list(gio.llr(np.asarray(logits), _allele_index(snpDF["ref"]), _allele_index(snpDF["alt"])))


def seq_from_vcf(args) -> Tuple[np.ndarray, List[int], list, list]:
    """Windows for every VCF record with an SNV ALT (reference :172-214).  Returns
    (ASCII [n, 512], record indices, header lines, records)."""
    logging.info(f"Reading input data from {args.inputVCF}")
    fasta = gio.read_fasta(args.inputFasta)
    header, records = gio.read_vcf(args.inputVCF)
    try:
        windows, record_indices = gio.windows_from_vcf(records, fasta, args.tokenIdx, 512)
    except KeyError as e:
        print(e.args[0])
        print("Check that VCF file is sorted and chromosome names match FASTA file.")
        raise SystemExit(1)
    return windows, record_indices, header, records


def zero_shot_score_vcf(args, recordIndices, logits, header, records):
    logging.info("Calculating zero-shot scores")
    gio.write_scored_vcf(args.output, header, records, recordIndices, logits)


def main(argv: Optional[Sequence[str]] = None):
    import pandas as pd
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    args = parse_args(argv)
    if args.inputDF is not None:
        logging.info(f"Reading input data from {args.inputDF}")
        snpDF = pd.read_csv(args.inputDF, delimiter="\t")
        logging.info("Filtering out invalid SNPs")
        valid = snpDF["ref"].isin(list(gio.NUCLEOTIDES)) & snpDF["alt"].isin(list(gio.NUCLEOTIDES))
        logging.info(f"Filtered out {len(snpDF) - int(valid.sum())} invalid SNPs")
        snpDF = snpDF[valid].copy()
        sequences = snpDF["sequences"].tolist()
    else:
        windows, recordIndices, header, records = seq_from_vcf(args)
        sequences = [bytes(r).decode("ascii") for r in windows]

    # One process per GPU under torchrun: contiguous window ranges per rank, scores gathered to every rank
    # (SURVEY.md 8e); a single process otherwise.
    import torch

    from . import sharding
    rank, local_rank, world = sharding.env_world()
    device = f"cuda:{local_rank}" if world > 1 else args.device
    if world > 1:
        torch.cuda.set_device(local_rank)
        sharding.init_process_group(device=torch.device(device))
    lo, hi = sharding.shard_range(len(sequences), rank, world)
    model, tokenizer = load_model_and_tokenizer(args.model, device, args.dtype, args.seed)
    logging.info("Creating data loader")
    loader = create_dataloader(sequences[lo:hi], tokenizer, args.batchSize, args.tokenIdx)
    logits = extract_logits(model, loader, device, args.tokenIdx, tokenizer)
    if world > 1:
        logits = sharding.gather_rows(torch.from_numpy(logits).to(device), len(sequences)).cpu().numpy()
        torch.distributed.barrier()
        if rank != 0:
            torch.distributed.destroy_process_group()
            return 0

    if args.inputDF is not None:
        snpDF["zeroShotScore"] = zero_shot_score(snpDF, logits)
        if args.outBED:
            logging.info("Outputting results in BED format")
            snpDF["start"] = snpDF["pos"] - 1
            snpDF["end"] = snpDF["pos"]
            snpDF[["chr", "start", "end", "ref", "alt", "zeroShotScore"]].to_csv(args.output, sep="\t", index=False, header=False)
        else:
            logging.info("Outputting results in tab-separated format")
            snpDF.to_csv(args.output, sep="\t", index=False)
    else:
        zero_shot_score_vcf(args, recordIndices, logits, header, records)
    logging.info(f"Zero-shot scores saved to {args.output}")
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
