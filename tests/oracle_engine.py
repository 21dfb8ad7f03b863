"""CPU stand-in for ``CaduceusForMaskedLM`` backed by the oracle (test infrastructure only): the engine's Python-facing
methods -- ``__call__`` / ``forward``, ``score_masked``, ``hidden_at``, ``extract_windows_device``, ``score_windows_device`` --
with the oracle's fp32 forward behind them, so that the host-side callers (CLI, zero_shot_eval, embeddings, genome_scan,
mutagenesis) can be checked WITHOUT a GPU against the outputs of the reference's own code (tests/golden/reference_run/).
The same checks run against the real engine under ``-m gpu``."""
from types import SimpleNamespace

import numpy as np
import torch

from oracle import caduceus_oracle as O
from plantcaduceus_b200 import CaduceusConfig, CharDNATokenizer, random_init_state_dict
from plantcaduceus_b200 import genome_io as gio


def tiny_config():
    """The model of tests/golden/make_reference_run_golden.py: 2 layers, d_model 128, random-init seed 0."""
    cfg = CaduceusConfig(d_model=128, n_layer=2)
    return cfg, random_init_state_dict(cfg, seed=0)


class OracleEngine:
    device = torch.device("cpu")
    dtype = torch.float32

    def __init__(self, cfg=None, sd=None, tokenizer=None):
        if cfg is None:
            cfg, sd = tiny_config()
        self.config, self._sd = cfg, sd
        self._tokenizer = tokenizer or CharDNATokenizer()
        v = self._tokenizer.get_vocab()
        self._acgt = [v[c] for c in "acgt"]

    def set_tokenizer(self, tok):
        self._tokenizer = tok

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def _run(self, ids, hidden=False):
        with torch.inference_mode():
            return O.caduceus_forward(self._sd, self.config, ids.long(), dtype=torch.float32, output_hidden_states=hidden)

    def forward(self, input_ids=None, output_hidden_states=False, compute_logits=True, **_kw):
        logits, hs = self._run(input_ids, output_hidden_states)
        return SimpleNamespace(logits=logits, hidden_states=(hs[-1],) if hs is not None else None)

    __call__ = forward

    def score_masked(self, ids_u8, positions, check_ids=True):
        logits, _ = self._run(ids_u8)
        if isinstance(positions, int):
            return logits[:, positions, self._acgt][:, None, :].clone()
        pos = positions.long()
        if pos.dim() == 1:
            pos = pos[:, None]
        rows = torch.arange(len(ids_u8))[:, None].expand_as(pos)
        return logits[rows, pos][..., self._acgt].clone()

    def hidden_at(self, ids_u8, positions):
        _, hs = self._run(ids_u8, True)
        pos = positions.long()
        if pos.dim() == 1:
            pos = pos[:, None]
        rows = torch.arange(len(ids_u8))[:, None].expand_as(pos)
        return hs[-1][rows, pos].clone()

    def extract_windows_device(self, chrom_dev, pos0, token_idx=255, length=512):
        chrom = bytes(chrom_dev.numpy())
        rows = [np.frombuffer(gio.extract_window(chrom, int(p), token_idx, length), dtype=np.uint8) for p in pos0.tolist()]
        return torch.from_numpy(np.stack(rows).copy()) if rows else torch.zeros((0, length), dtype=torch.uint8)

    def score_windows_device(self, ascii_dev, token_idx, out=None):
        ids = torch.from_numpy(self._tokenizer.encode_bytes(ascii_dev.numpy()).copy())
        if len(ids):
            ids[:, token_idx] = self._tokenizer.mask_token_id
            res = self.score_masked(ids, int(token_idx))[:, 0]
        else:
            res = torch.zeros((0, 4))
        return res if out is None else out.copy_(res)


def oracle_extract_logits(model, dataloader, device, tokenIdx, tokenizer):
    """``zero_shot_score.extract_logits`` without the pinned-memory / CUDA-event double buffering: same batches, same call."""
    dataset, batch_size = dataloader
    out = np.zeros((len(dataset), 4), dtype=np.float32)
    for start, batch in dataset.ascii_batches(batch_size):
        L = batch.shape[1]
        out[start:start + len(batch)] = model.score_windows_device(torch.from_numpy(np.ascontiguousarray(batch)), tokenIdx % L).numpy()
    return gio.softmax4(out) if len(dataset) else np.zeros((0, 4), dtype=np.float32)
