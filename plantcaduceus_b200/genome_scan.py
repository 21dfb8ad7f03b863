"""Genome-scale zero-shot SNP scoring with the chromosome resident in HBM (BASELINE.json config 3: millions of
variants sharded by window across the GPUs of one box).

The reference builds every window on the host (`seq_from_vcf`, src/zero_shot_score.py:172-214: Biopython dict of the
whole genome, one Python slice per record) and tokenises it per sequence.  Here a chromosome is copied to the device
once; each batch sends only the variant positions (8 bytes per variant), and window extraction
(`pcad_extract_windows`, same slice-and-pad rule), tokenisation, masking, the forward pass and the LM head at the
masked index all run on the device.  Variants are sharded contiguously over ranks (sharding.py) and the per-variant
logits gathered ONCE at the end.  Only rank 0 touches the input files: the other ranks receive the variant coordinates
and each chromosome's bytes over the process group (NCCL broadcast GPU to GPU).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import genome_io as gio
from . import sharding


def score_positions(model, chrom: bytes, pos0: np.ndarray, batch_size: int = 256, token_idx: int = 255,
                    length: int = 512, chrom_dev: Optional[torch.Tensor] = None) -> np.ndarray:
    """softmax(a,c,g,t) at the masked variant position for every 0-based position in ``pos0`` of ``chrom``:
    float32 [n, 4], this rank's share computed locally and the full result gathered on every rank."""
    rank, _local, world = sharding.env_world()
    pos0 = np.asarray(pos0, dtype=np.int64)
    n = len(pos0)
    lo, hi = sharding.shard_range(n, rank, world)
    if chrom_dev is None:
        chrom_dev = torch.from_numpy(np.frombuffer(chrom, dtype=np.uint8).copy()).to(model.device)
    local = score_positions_local(model, chrom_dev, torch.from_numpy(pos0[lo:hi]).to(model.device), batch_size, token_idx, length)
    full = sharding.gather_rows(local, n) if world > 1 else local
    return gio.softmax4(full.cpu().numpy())


def score_positions_local(model, chrom_dev: torch.Tensor, pos_dev: torch.Tensor, batch_size: int = 256, token_idx: int = 255,
                          length: int = 512, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The per-rank inner loop: raw a,c,g,t logits [m, 4] (device, asynchronous) for the positions ``pos_dev`` (int64, on
    the device) of the resident chromosome.  No host synchronisation, no collective."""
    m = int(pos_dev.numel())
    if out is None:
        out = torch.empty((m, 4), dtype=torch.float32, device=pos_dev.device)
    for s in range(0, m, batch_size):
        e = min(s + batch_size, m)
        windows = model.extract_windows_device(chrom_dev, pos_dev[s:e], token_idx, length)
        model.score_windows_device(windows, token_idx, out=out[s:e])
    return out


def score_variants_sharded(model, chrom_names: Optional[Sequence[str]], chrom_seqs: Optional[Dict[str, bytes]],
                           chrom_id: Optional[np.ndarray], pos0: Optional[np.ndarray], batch_size: int = 256,
                           token_idx: int = 255, length: int = 512, device: Optional[torch.device] = None) -> np.ndarray:
    """Raw a,c,g,t logits float32 [n, 4] for variants on several chromosomes, in input order, on every rank.

    Rank 0 supplies everything (``chrom_names``; ``chrom_seqs`` name -> bytes; per-variant ``chrom_id`` into
    ``chrom_names`` and 0-based ``pos0``); the other ranks pass ``None`` and receive what they need: the coordinates
    (12 bytes per variant) by broadcast, and each chromosome that carries variants as one device-to-device broadcast,
    dropped again before the next one arrives (peak = one chromosome in HBM, as the reference frees finished chromosomes,
    zero_shot_score.py:204-207).  Each rank scores its contiguous share of the variant list."""
    rank, _local, world = sharding.env_world()
    device = torch.device(device) if device is not None else model.device
    multi = world > 1 and dist.is_initialized()
    meta = [None]
    if rank == 0:
        chrom_id = np.ascontiguousarray(chrom_id, dtype=np.int32)
        pos0 = np.ascontiguousarray(pos0, dtype=np.int64)
        meta = [(list(chrom_names), [len(chrom_seqs[c]) for c in chrom_names], len(pos0))]
    if multi:
        dist.broadcast_object_list(meta, src=0)
    names, lens, n = meta[0]
    cid_t = torch.from_numpy(chrom_id).to(device) if rank == 0 else torch.empty(n, dtype=torch.int32, device=device)
    pos_t = torch.from_numpy(pos0).to(device) if rank == 0 else torch.empty(n, dtype=torch.int64, device=device)
    if multi and n:
        dist.broadcast(cid_t, src=0)
        dist.broadcast(pos_t, src=0)
    lo, hi = sharding.shard_range(n, rank, world)
    local = torch.zeros((hi - lo, 4), dtype=torch.float32, device=device)
    used = torch.unique(cid_t).cpu().tolist() if n else []
    my_cid, my_pos = cid_t[lo:hi], pos_t[lo:hi]
    for c in used:
        buf = torch.empty(lens[c], dtype=torch.uint8, device=device)
        if rank == 0:
            buf.copy_(torch.from_numpy(np.frombuffer(chrom_seqs[names[c]], dtype=np.uint8).copy()))
        if multi:
            dist.broadcast(buf, src=0)
        sel = torch.nonzero(my_cid == c, as_tuple=False).flatten()
        if sel.numel():
            part = score_positions_local(model, buf, my_pos[sel], batch_size, token_idx, length)
            local[sel] = part
        del buf
    full = sharding.gather_rows(local, n) if multi else local
    return full.cpu().numpy()
