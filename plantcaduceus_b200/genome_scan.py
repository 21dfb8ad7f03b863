"""Genome-scale zero-shot SNP scoring with the chromosome resident in HBM (BASELINE.json config 3: millions of
variants sharded by window across the GPUs of one box).

The reference builds every window on the host (`seq_from_vcf`, src/zero_shot_score.py:172-214: Biopython dict of the
whole genome, one Python slice per record) and tokenises it per sequence.  Here the chromosome is copied to the device
once; each batch sends only the variant positions (8 bytes per variant), and window extraction
(`pcad_extract_windows`, same slice-and-pad rule), tokenisation, masking, the forward pass and the LM head at the
masked index all run on the device.  Variants are sharded contiguously over ranks (sharding.py) and the per-variant
logits gathered at the end.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

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
    local = torch.empty((hi - lo, 4), dtype=torch.float32, device=model.device)
    pos_dev = torch.from_numpy(pos0[lo:hi]).to(model.device)
    for s in range(0, hi - lo, batch_size):
        e = min(s + batch_size, hi - lo)
        windows = model.extract_windows_device(chrom_dev, pos_dev[s:e], token_idx, length)
        local[s:e] = model.score_windows_device(windows, token_idx)
    full = sharding.gather_rows(local, n) if world > 1 else local
    return gio.softmax4(full.cpu().numpy())
