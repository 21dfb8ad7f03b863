"""State-dict naming and seeded random initialisation for Caduceus weights.

Names are the ones the HF-hub checkpoint carries (module tree printed at reference
notebooks/examples.ipynb:61-98; Mamba parameter names corroborated by the LoRA targets at
reference src/lora_fine_tune.py:609-617):

    caduceus.backbone.embeddings.word_embeddings.embedding.weight            [V, d]
    caduceus.backbone.layers.{i}.mixer.submodule.mamba_{fwd,rev}.in_proj.weight   [2E, d]   (rev aliases fwd)
    ...                                                   .conv1d.weight     [E, 1, 4]
    ...                                                   .conv1d.bias       [E]
    ...                                                   .x_proj.weight     [R+2N, E]
    ...                                                   .dt_proj.weight    [E, R]
    ...                                                   .dt_proj.bias      [E]
    ...                                                   .A_log             [E, N]
    ...                                                   .D                 [E]
    ...                                                   .out_proj.weight   [d, E]   (rev aliases fwd)
    caduceus.backbone.layers.{i}.norm.weight                                 [d]
    caduceus.backbone.norm_f.weight                                          [d]
    lm_head.lm_head.weight                                                   [V, d]   (tied to the embedding)

There is no network in the build/bench environment, so benchmarks and parity tests run on
random-initialised weights of the published architecture (BASELINE.json configs 1-2).  The
initialisation below follows the Mamba-1 recipe in spirit (S4D-real A with jitter so that the
A[e,n] = -(n+1) structure cannot be exploited by a kernel, inverse-softplus dt bias drawn
log-uniformly in [1e-3, 1e-1]) and is fully determined by ``seed`` through a CPU generator.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .configuration import CaduceusConfig

PREFIX = "caduceus.backbone."
EMB_KEY = PREFIX + "embeddings.word_embeddings.embedding.weight"
HEAD_KEY = "lm_head.lm_head.weight"
NORM_F_KEY = PREFIX + "norm_f.weight"
DIRS = ("mamba_fwd", "mamba_rev")
PER_DIR = ("conv1d.weight", "conv1d.bias", "x_proj.weight", "dt_proj.weight", "dt_proj.bias", "A_log", "D")
# Mamba-2 ([EXT] mamba_ssm.modules.mamba2.Mamba2): in_proj [2E + 2GN + H, d] and out_proj [d, E] tied between the directions;
# per direction conv1d.weight [E + 2GN, 1, 4], conv1d.bias [E + 2GN], dt_bias [H], A_log [H], D [H], norm.weight [E]
PER_DIR_M2 = ("conv1d.weight", "conv1d.bias", "dt_bias", "A_log", "D", "norm.weight")
SHARED = ("in_proj.weight", "out_proj.weight")


def layer_key(i: int, direction: str, name: str) -> str:
    return f"{PREFIX}layers.{i}.mixer.submodule.{direction}.{name}"


def norm_key(i: int) -> str:
    return f"{PREFIX}layers.{i}.norm.weight"


def random_init_state_dict(cfg: CaduceusConfig, seed: int = 0,
                           dtype: torch.dtype = torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random weights with the checkpoint's names, shapes and tying."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    d, E, N, R, K, V = cfg.d_model, cfg.d_inner, cfg.d_state, cfg.dt_rank, cfg.d_conv, cfg.vocab_size

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound

    def normal(shape, std):
        return torch.randn(shape, generator=g, dtype=torch.float32) * std

    sd: Dict[str, torch.Tensor] = {}
    emb = normal((V, d), 0.02)
    sd[EMB_KEY] = emb
    sd[HEAD_KEY] = emb  # tied (225.36 M parameter count)
    if cfg.is_mamba2:
        H, CD, DIP = cfg.nheads, cfg.conv_dim, cfg.d_in_proj
        for i in range(cfg.n_layer):
            w_in = uniform((DIP, d), 1.0 / math.sqrt(d))
            w_out = uniform((d, E), 1.0 / math.sqrt(E)) / math.sqrt(2.0 * cfg.n_layer)
            for direction in DIRS:
                sd[layer_key(i, direction, "in_proj.weight")] = w_in
                sd[layer_key(i, direction, "out_proj.weight")] = w_out
                sd[layer_key(i, direction, "conv1d.weight")] = uniform((CD, 1, K), 1.0 / math.sqrt(K))
                sd[layer_key(i, direction, "conv1d.bias")] = uniform((CD,), 1.0 / math.sqrt(K))
                dt = torch.exp(torch.rand((H,), generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3)).clamp(min=1e-4)
                sd[layer_key(i, direction, "dt_bias")] = dt + torch.log(-torch.expm1(-dt))
                sd[layer_key(i, direction, "A_log")] = torch.log(1.0 + 15.0 * torch.rand((H,), generator=g))   # A in (1, 16)
                sd[layer_key(i, direction, "D")] = 1.0 + normal((H,), 0.1)
                sd[layer_key(i, direction, "norm.weight")] = 1.0 + normal((E,), 0.1)
            sd[norm_key(i)] = 1.0 + normal((d,), 0.1)
    for i in range(0 if cfg.is_mamba2 else cfg.n_layer):
        w_in = uniform((2 * E, d), 1.0 / math.sqrt(d))
        w_out = uniform((d, E), 1.0 / math.sqrt(E)) / math.sqrt(2.0 * cfg.n_layer)
        for direction in DIRS:
            sd[layer_key(i, direction, "in_proj.weight")] = w_in
            sd[layer_key(i, direction, "out_proj.weight")] = w_out
            sd[layer_key(i, direction, "conv1d.weight")] = uniform((E, 1, K), 1.0 / math.sqrt(K))
            sd[layer_key(i, direction, "conv1d.bias")] = uniform((E,), 1.0 / math.sqrt(K))
            sd[layer_key(i, direction, "x_proj.weight")] = uniform((R + 2 * N, E), 1.0 / math.sqrt(E))
            sd[layer_key(i, direction, "dt_proj.weight")] = uniform((E, R), R ** -0.5)
            dt = torch.exp(torch.rand((E,), generator=g) * (math.log(0.1) - math.log(1e-3)) + math.log(1e-3))
            dt = dt.clamp(min=1e-4)
            sd[layer_key(i, direction, "dt_proj.bias")] = dt + torch.log(-torch.expm1(-dt))
            a = torch.arange(1, N + 1, dtype=torch.float32).repeat(E, 1)
            sd[layer_key(i, direction, "A_log")] = torch.log(a) + normal((E, N), 0.1)
            sd[layer_key(i, direction, "D")] = 1.0 + normal((E,), 0.1)
        sd[norm_key(i)] = 1.0 + normal((d,), 0.1)
    sd[NORM_F_KEY] = 1.0 + normal((d,), 0.1)
    if dtype != torch.float32:
        # from_pretrained(torch_dtype=...) casts every floating parameter, A_log and D included.
        cast = {k: v.to(dtype) for k, v in sd.items()}
        cast[HEAD_KEY] = cast[EMB_KEY]
        sd = cast
    return sd


def count_parameters(sd: Dict[str, torch.Tensor]) -> int:
    """Unique parameter count (tied tensors counted once), for the README.md:60-63 check."""
    seen, total = set(), 0
    for v in sd.values():
        p = v.data_ptr()
        if p in seen:
            continue
        seen.add(p)
        total += v.numel()
    return total


def restore_tied_names(sd: Dict[str, torch.Tensor], n_layer: int) -> Dict[str, torch.Tensor]:
    """A checkpoint saved with tied tensors de-duplicated carries one name per tie; after loading, HF's ``tie_weights``
    makes both names resolve to the same tensor again.  Does the same in place: embedding <-> LM head, and
    ``mamba_fwd`` <-> ``mamba_rev`` for ``in_proj.weight`` / ``out_proj.weight`` (bidirectional_weight_tie)."""
    for a, b in ((EMB_KEY, HEAD_KEY),):
        if a not in sd and b in sd:
            sd[a] = sd[b]
        elif b not in sd and a in sd:
            sd[b] = sd[a]
    for i in range(n_layer):
        for leaf in SHARED:
            f, r = layer_key(i, "mamba_fwd", leaf), layer_key(i, "mamba_rev", leaf)
            if f not in sd and r in sd:
                sd[f] = sd[r]
            elif r not in sd and f in sd:
                sd[r] = sd[f]
    return sd


def write_checkpoint_dir(path: str, cfg: CaduceusConfig, sd: Dict[str, torch.Tensor], vocab=None) -> None:
    """Writes a directory laid out like an HF-hub Caduceus snapshot (what ``save_pretrained`` leaves and what the
    reference loads at src/zero_shot_score.py:91,96): ``config.json`` (string-keyed ``complement_map``, nested
    ``ssm_cfg``), ``model.safetensors`` with tied tensors de-duplicated and the RCPS modules' int64 ``complement_map``
    buffers, ``tokenizer.json`` / ``tokenizer_config.json``.  Used by the checkpoint-directory tests and by
    tools/readme_benchmark.py (no hub checkpoint is reachable offline)."""
    import json
    import os

    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    d = cfg.to_dict()
    d["architectures"] = ["CaduceusForMaskedLM"]
    d["auto_map"] = {"AutoConfig": "configuration_caduceus.CaduceusConfig",
                     "AutoModelForMaskedLM": "modeling_caduceus.CaduceusForMaskedLM"}
    d["torch_dtype"] = "float32"
    d["transformers_version"] = "4.40.0"
    d["initializer_cfg"] = {"initializer_range": 0.02, "rescale_prenorm_residual": True, "n_residuals_per_layer": 1}
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(d, f)
    out = {}
    comp = torch.tensor([cfg.complement_map[i] for i in range(cfg.vocab_size)], dtype=torch.int64)
    out["caduceus.backbone.embeddings.word_embeddings.complement_map"] = comp.clone()
    out["lm_head.complement_map"] = comp.clone()
    for k, v in sd.items():
        if k == HEAD_KEY:
            continue                       # tied to the embedding: de-duplicated on save
        out[k] = v.clone().contiguous()
    for i in range(cfg.n_layer):           # tied in/out projections: only one of the two names survives de-duplication
        drop = "mamba_rev" if i % 2 == 0 else "mamba_fwd"
        for leaf in SHARED:
            del out[layer_key(i, drop, leaf)]
    save_file(out, os.path.join(path, "model.safetensors"))
    vocab = vocab or {"[PAD]": 0, "[MASK]": 1, "[UNK]": 2, "a": 3, "c": 4, "g": 5, "t": 6}
    with open(os.path.join(path, "tokenizer.json"), "w") as f:
        json.dump({"version": "1.0", "model": {"type": "WordLevel", "vocab": vocab, "unk_token": "[UNK]"},
                   "normalizer": {"type": "Lowercase"}}, f)
    with open(os.path.join(path, "tokenizer_config.json"), "w") as f:
        json.dump({"mask_token": "[MASK]", "pad_token": "[PAD]", "unk_token": "[UNK]"}, f)
