"""Checkpoint-directory ingest: what ``AutoModelForMaskedLM.from_pretrained(model_dir, trust_remote_code=True)`` and
``AutoTokenizer.from_pretrained(model_dir)`` read at reference src/zero_shot_score.py:91,96 -- ``config.json`` (JSON object keys are
strings, so ``complement_map`` arrives string-keyed; ``ssm_cfg`` nested), ``model.safetensors`` as HF saves it (tied tensors
de-duplicated, the RCPS modules' int64 ``complement_map`` buffers present) and ``tokenizer.json``."""
import json
import os

import numpy as np
import pytest
import torch

from plantcaduceus_b200 import CaduceusConfig, CharDNATokenizer, random_init_state_dict
from plantcaduceus_b200.weights import EMB_KEY, HEAD_KEY, layer_key


def write_checkpoint_dir(path, cfg, sd, vocab=None):
    """A directory laid out like an HF-hub Caduceus snapshot (plantcaduceus_b200.weights.write_checkpoint_dir)."""
    from plantcaduceus_b200.weights import write_checkpoint_dir as w
    w(path, cfg, sd, vocab)
    assert all(isinstance(k, str) for k in json.load(open(os.path.join(path, "config.json")))["complement_map"])
    from safetensors import safe_open
    with safe_open(os.path.join(path, "model.safetensors"), "pt") as f:
        keys = set(f.keys())
    assert HEAD_KEY not in keys and "lm_head.complement_map" in keys
    assert layer_key(0, "mamba_rev", "in_proj.weight") not in keys and layer_key(0, "mamba_fwd", "in_proj.weight") in keys


def test_checkpoint_dir_parses_on_cpu(tmp_path):
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    cfg = CaduceusConfig(d_model=128, n_layer=2, norm_epsilon=1e-6)
    sd = random_init_state_dict(cfg, seed=3)
    write_checkpoint_dir(str(tmp_path / "ckpt"), cfg, sd)
    m = CaduceusForMaskedLM.from_pretrained(str(tmp_path / "ckpt"), trust_remote_code=True, torch_dtype=torch.float32)
    assert m.config.d_model == 128 and m.config.n_layer == 2 and m.config.norm_epsilon == 1e-6
    assert m.config.complement_map == cfg.complement_map and all(isinstance(k, int) for k in m.config.complement_map)
    assert m.config.d_state == 16 and m.config.dt_rank == 8
    assert EMB_KEY in m.state_dict()
    tok = CharDNATokenizer.from_pretrained(str(tmp_path / "ckpt"))
    assert tok.get_vocab()["a"] == 3 and tok.mask_token_id == 1


def test_validate_supported_mirrors_engine_limits():
    for kw in (dict(d_model=192), dict(d_model=4096), dict(fused_add_norm=False),
               dict(ssm_cfg=dict(d_state=16, d_conv=4, expand=4, dt_rank="auto", conv_bias=True, bias=False)),
               dict(ssm_cfg=dict(d_state=16, d_conv=4, expand=2, dt_rank=12, conv_bias=True, bias=False))):
        with pytest.raises(ValueError, match="unsupported Caduceus configuration"):
            CaduceusConfig(**{"d_model": 128, "n_layer": 1, **kw}).validate_supported()
    CaduceusConfig(d_model=128, n_layer=1).validate_supported()


@pytest.mark.gpu
def test_checkpoint_dir_forward_matches_oracle(tmp_path, cuda_device):
    from oracle import caduceus_oracle as O
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    cfg = CaduceusConfig(d_model=256, n_layer=3)
    sd = random_init_state_dict(cfg, seed=11)
    write_checkpoint_dir(str(tmp_path / "ckpt"), cfg, sd)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(3, 7, (3, 128), generator=g)
    ids[:, 64] = 1
    want, _ = O.caduceus_forward(sd, cfg, ids, dtype=torch.float32)
    m = CaduceusForMaskedLM.from_pretrained(str(tmp_path / "ckpt"), trust_remote_code=True,
                                            torch_dtype=torch.float32).to(cuda_device)
    got = m(input_ids=ids.to(cuda_device)).logits.cpu()
    assert ((got - want).abs().max() / want.abs().max()).item() <= 1e-4


@pytest.mark.gpu
def test_checkpoint_dir_with_permuted_vocab_through_cli(tmp_path, cuda_device):
    """A tokenizer whose ids are NOT the default order: the engine must take ids, mask id, a/c/g/t columns and the
    complement map from the directory, not from its defaults.  Scores through the CLI's -model <dir> equal the oracle's."""
    import pandas as pd
    from oracle import caduceus_oracle as O
    from plantcaduceus_b200 import build_complement_map, zero_shot_score as zs
    vocab = {"[PAD]": 0, "[UNK]": 1, "[MASK]": 2, "t": 3, "a": 4, "g": 5, "c": 6}
    cfg = CaduceusConfig(d_model=128, n_layer=2, complement_map=build_complement_map(vocab, 8))
    assert cfg.complement_map[3] == 4 and cfg.complement_map[5] == 6
    sd = random_init_state_dict(cfg, seed=5)
    ckpt = str(tmp_path / "ckpt")
    write_checkpoint_dir(ckpt, cfg, sd, vocab=vocab)
    src = os.path.join(os.path.dirname(__file__), "golden", "example_snp.tsv")
    df = pd.read_csv(src, sep="\t").head(12)
    table = str(tmp_path / "in.tsv")
    df.to_csv(table, sep="\t", index=False)
    out = str(tmp_path / "out.tsv")
    assert zs.main(["-input-table", table, "-output", out, "-model", ckpt, "-dtype", "float32", "-batchSize", "5"]) == 0
    got = pd.read_csv(out, sep="\t")
    tok = CharDNATokenizer(vocab=vocab)
    keep = df[df["ref"].isin(list("ACGT")) & df["alt"].isin(list("ACGT"))]
    ids = torch.cat([tok.encode_plus(s, return_tensors="pt")["input_ids"] for s in keep["sequences"]])
    ids[:, 255] = tok.mask_token_id
    logits, _ = O.caduceus_forward(sd, cfg, ids, dtype=torch.float32)
    probs = O.extract_acgt_probs(logits, 255, [vocab[c] for c in "acgt"])
    want = O.zero_shot_llr(probs, list(keep["ref"]), list(keep["alt"]))
    assert len(got) == len(want)
    assert np.allclose(got["zeroShotScore"].to_numpy(), np.array(want), rtol=1e-3, atol=1e-4)


@pytest.mark.gpu
def test_out_of_range_ids_raise(cuda_device):
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    m = CaduceusForMaskedLM.from_random(CaduceusConfig(d_model=128, n_layer=1), seed=0).to(cuda_device)
    ids = torch.full((2, 16), 3, dtype=torch.long)
    ids[1, 5] = 8
    with pytest.raises(IndexError, match="token id outside"):
        m(input_ids=ids.to(cuda_device))
    m(input_ids=torch.full((2, 16), 3, dtype=torch.long, device=cuda_device))      # the flag was cleared
    u8 = torch.full((2, 16), 4, dtype=torch.uint8)
    u8[0, 0] = 200
    with pytest.raises(IndexError):
        m.score_masked(u8, torch.zeros((2, 1), dtype=torch.int32))
    ok = m.score_masked(torch.full((2, 16), 4, dtype=torch.uint8), torch.zeros((2, 1), dtype=torch.int32))
    assert torch.isfinite(ok).all()


def test_tokenizer_json_variants(tmp_path):
    """tokenizer.json as the `tokenizers` library writes it: special tokens possibly only under ``added_tokens``; a
    list-style vocabulary."""
    d = tmp_path / "a"
    d.mkdir()
    (d / "tokenizer.json").write_text(json.dumps({
        "added_tokens": [{"id": 0, "content": "[PAD]", "special": True}, {"id": 1, "content": "[MASK]", "special": True},
                         {"id": 2, "content": "[UNK]", "special": True}],
        "model": {"type": "WordLevel", "vocab": {"a": 3, "c": 4, "g": 5, "t": 6}, "unk_token": "[UNK]"}}))
    tok = CharDNATokenizer.from_pretrained(str(d))
    assert (tok.pad_token_id, tok.mask_token_id, tok.unk_token_id) == (0, 1, 2) and tok.encode("ACGTN") == [3, 4, 5, 6, 2]
    e = tmp_path / "b"
    e.mkdir()
    (e / "tokenizer.json").write_text(json.dumps({"model": {"type": "Unigram", "vocab": [["[PAD]", 0.0], ["[MASK]", 0.0], ["[UNK]", 0.0],
                                                                                          ["t", -1.0], ["g", -1.0], ["c", -1.0], ["a", -1.0]]}}))
    tok = CharDNATokenizer.from_pretrained(str(e))
    assert tok.encode("acgt") == [6, 5, 4, 3] and tok.mask_token_id == 1
