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
    scorer itself consumes whole ``ascii_batch`` slices and tokenises on the device.  ``sequences`` is a list of
    strings (possibly ragged, as the reference's table may be) or an already packed uint8 ``[n, L]`` matrix."""

    def __init__(self, sequences, tokenizer: CharDNATokenizer, tokenIdx: int):
        self.matrix = sequences if isinstance(sequences, np.ndarray) else None
        self.sequences = None if self.matrix is not None else list(sequences)
        self.tokenizer = tokenizer
        self.tokenIdx = tokenIdx

    def __len__(self):
        return len(self.matrix) if self.matrix is not None else len(self.sequences)

    def __getitem__(self, idx):
        seq = bytes(self.matrix[idx]).decode("latin-1") if self.matrix is not None else self.sequences[idx]
        ids = self.tokenizer.encode_plus(seq, return_tensors="pt", return_attention_mask=False,
                                         return_token_type_ids=False)["input_ids"]
        ids[0, self.tokenIdx] = self.tokenizer.mask_token_id
        return {"sequence": seq, "input_ids": ids}

    def ascii_batches(self, batch_size: int):
        """Yields (start, uint8 [b, L]) over runs of equal-length windows, in input order."""
        if self.matrix is not None:
            for i in range(0, len(self.matrix), batch_size):
                yield i, self.matrix[i:i + batch_size]
            return
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
    (reference extract_logits, :107-121).  Batches are double-buffered: while the GPU scores batch i from one pinned
    staging buffer, the host packs batch i+1 into the other; results stay on the device until one copy at the end (the
    reference synchronises and copies per batch, :119)."""
    import torch
    dataset, batch_size = dataloader
    logging.info("Extracting logits")
    n = len(dataset)
    dev = model.device
    logits = torch.zeros((n, 4), dtype=torch.float32, device=dev)
    ring = [None, None]          # (pinned host buffer, device buffer, event recorded after the H2D copy was consumed)
    for k, (start, ascii_batch) in enumerate(dataset.ascii_batches(batch_size)):
        b, L = ascii_batch.shape
        if tokenIdx >= L or tokenIdx < -L:
            raise IndexError(f"tokenIdx {tokenIdx} is out of bounds for a window of length {L}")
        slot = ring[k & 1]
        if slot is None or slot[0].shape[0] < b or slot[0].shape[1] != L:
            slot = (torch.empty((max(b, batch_size), L), dtype=torch.uint8).pin_memory(),
                    torch.empty((max(b, batch_size), L), dtype=torch.uint8, device=dev), torch.cuda.Event())
        else:
            slot[2].synchronize()      # the copy that last used this pinned buffer has completed
        host, devbuf, ev = slot
        host[:b].copy_(torch.from_numpy(ascii_batch))
        with torch.cuda.device(dev):
            devbuf[:b].copy_(host[:b], non_blocking=True)
            ev.record()
            model.score_windows_device(devbuf[:b], tokenIdx % L, out=logits[start:start + b])
        ring[k & 1] = slot
    return gio.softmax4(logits.cpu().numpy()) if n else np.zeros((0, 4), dtype=np.float32)


def _allele_index(values) -> np.ndarray:
    """Index into A,C,G,T of every (single upper-case letter) allele; anything else is a KeyError, as in the
    reference's ``nucleotides.index`` / column lookup (:127-133)."""
    values = list(values)
    flat = np.frombuffer("".join(values).encode("latin-1", errors="replace"), dtype=np.uint8)
    if len(flat) != len(values):
        raise KeyError(next(v for v in values if len(v) != 1))
    lut = np.full(256, -1, dtype=np.int64)
    lut[[ord(c) for c in gio.NUCLEOTIDES]] = np.arange(4)
    idx = lut[flat]
    if (idx < 0).any():
        raise KeyError(values[int(np.flatnonzero(idx < 0)[0])])
    return idx


def zero_shot_score(snpDF, logits) -> List[float]:
    """log(p_alt / p_ref) per row of the SNP table (reference :124-134)."""
    logging.info("Calculating zero-shot scores")
    return list(gio.llr(np.asarray(logits), _allele_index(snpDF["ref"]), _allele_index(snpDF["alt"])))


def seq_from_vcf(args) -> Tuple[np.ndarray, List[int], list, list]:
    """Windows for every VCF record with an SNV ALT (reference :172-214), built on the host.  Returns
    (ASCII [n, 512], record indices, header lines, records).  ``main`` uses ``variants_from_vcf`` + device-side
    extraction instead; this keeps the reference's function for callers that want the windows themselves."""
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


def variants_from_vcf(args):
    """Rank 0's parse of the VCF + FASTA (reference :172-214) into coordinates instead of windows: returns
    (chrom names, {name: bytes}, chrom_id int32 [n], pos0 int64 [n], record indices int64 [n], header, VcfTable) for the
    records that carry at least one SNV ALT.  The VCF is parsed column-wise (``genome_io.read_vcf_table``: seconds for
    10 M records); the windows are cut on the device from the resident chromosome with the same slice-and-pad rule
    (pcad_extract_windows)."""
    logging.info(f"Reading input data from {args.inputVCF}")
    fasta = gio.read_fasta(args.inputFasta)
    table = gio.read_vcf_table(args.inputVCF)
    record_indices = np.flatnonzero(table.has_snv).astype(np.int64)
    cid = table.chrom_id[record_indices]
    # chromosomes that carry scored records, in order of first appearance
    first_seen = {}
    if len(cid):
        uniq, first = np.unique(cid, return_index=True)
        first_seen = {int(u): int(f) for u, f in zip(uniq, first)}
    used = sorted(first_seen, key=first_seen.get)
    for c in used:
        if table.chrom_names[c] not in fasta:
            print(f"VCF record {int(record_indices[first_seen[c]])}: chromosome {table.chrom_names[c]!r} is not in the FASTA "
                  f"(check that chromosome names match)")
            print("Check that VCF file is sorted and chromosome names match FASTA file.")
            raise SystemExit(1)
    remap = np.full(max(len(table.chrom_names), 1), -1, dtype=np.int32)
    remap[used] = np.arange(len(used), dtype=np.int32)
    names = [table.chrom_names[c] for c in used]
    seqs = {c: fasta[c] for c in names}
    chrom_id = remap[cid] if len(cid) else np.zeros(0, dtype=np.int32)
    pos0 = (table.pos[record_indices] - 1).astype(np.int64)
    return names, seqs, chrom_id.astype(np.int32), pos0, record_indices, table.header, table


def zero_shot_score_vcf(args, recordIndices, logits, header, records):
    """Writes the scored VCF (reference :137-169).  ``records`` is the ``VcfTable`` of ``variants_from_vcf`` or the
    ``VcfRecord`` list of ``seq_from_vcf``."""
    logging.info("Calculating zero-shot scores")
    if isinstance(records, gio.VcfTable):
        gio.write_scored_vcf_table(args.output, records, recordIndices, logits)
    else:
        gio.write_scored_vcf(args.output, header, records, recordIndices, logits)


def main(argv: Optional[Sequence[str]] = None):
    import pandas as pd
    import torch

    from . import genome_scan, sharding
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
    args = parse_args(argv)
    # One process per GPU under torchrun: contiguous variant ranges per rank, scores gathered once at the end
    # (SURVEY.md 8e); a single process otherwise.  Only rank 0 reads the input files.
    rank, local_rank, world = sharding.env_world()
    # "-device cpu" exists for the host-logic tests only (gloo ranks, stubbed model): the engine itself has no CPU path
    on_cpu = torch.device(args.device).type == "cpu"
    device = args.device if (world == 1 or on_cpu) else f"cuda:{local_rank}"
    if world > 1:
        if not on_cpu:
            torch.cuda.set_device(local_rank)
        sharding.init_process_group(device=torch.device(device))
    model, tokenizer = load_model_and_tokenizer(args.model, device, args.dtype, args.seed)

    snpDF = None
    if args.inputDF is not None:
        windows = None
        ragged = [None]
        if rank == 0:
            logging.info(f"Reading input data from {args.inputDF}")
            snpDF = pd.read_csv(args.inputDF, delimiter="\t")
            logging.info("Filtering out invalid SNPs")
            valid = snpDF["ref"].isin(list(gio.NUCLEOTIDES)) & snpDF["alt"].isin(list(gio.NUCLEOTIDES))
            logging.info(f"Filtered out {len(snpDF) - int(valid.sum())} invalid SNPs")
            snpDF = snpDF[valid].copy()
            sequences = snpDF["sequences"].tolist()
            if len({len(q) for q in sequences}) <= 1:
                windows = torch.from_numpy(tokenizer.windows_to_ascii(sequences, len(sequences[0]) if sequences else 512))
            else:      # ragged table (the reference never checks lengths): every rank gets the strings
                ragged = [sequences]
        if world > 1:
            torch.distributed.broadcast_object_list(ragged, src=0)
        if ragged[0] is not None:
            lo, hi = sharding.shard_range(len(ragged[0]), rank, world)
            mine, n_total = ragged[0][lo:hi], len(ragged[0])
        else:
            shard = sharding.scatter_rows(windows, device=torch.device(device)) if world > 1 else windows
            mine = np.ascontiguousarray(shard.cpu().numpy())
            n_meta = [len(windows) if rank == 0 else None]
            if world > 1:
                torch.distributed.broadcast_object_list(n_meta, src=0)
            n_total = n_meta[0]
        logging.info("Creating data loader")
        loader = create_dataloader(mine, tokenizer, args.batchSize, args.tokenIdx)
        logits = extract_logits(model, loader, device, args.tokenIdx, tokenizer)
        if world > 1:
            logits = sharding.gather_rows(torch.from_numpy(logits).to(device), n_total).cpu().numpy()
    else:
        parsed = variants_from_vcf(args) if rank == 0 else (None,) * 7
        names, seqs, chrom_id, pos0, recordIndices, header, records = parsed
        raw = genome_scan.score_variants_sharded(model, names, seqs, chrom_id, pos0, args.batchSize, args.tokenIdx, 512,
                                                 device=torch.device(device))
        logits = gio.softmax4(raw) if len(raw) else np.zeros((0, 4), dtype=np.float32)

    if world > 1:
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
