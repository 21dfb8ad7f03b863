"""BASELINE.json configs 4 and 5 as parity cases against the CPU oracle (small model sizes so the oracle finishes):
config 4 = saturation mutagenesis (every position masked in turn, 3 alts each); config 5 = long context (L = 8192,
512 scan chunks carried exactly) with the hidden-state tap the embedding extraction reads."""
import numpy as np
import pytest
import torch

from oracle import caduceus_oracle as O
from plantcaduceus_b200 import CaduceusConfig, random_init_state_dict

pytestmark = pytest.mark.gpu


def test_saturation_mutagenesis_matches_oracle(cuda_device):
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    from plantcaduceus_b200.mutagenesis import saturation_mutagenesis
    cfg = CaduceusConfig(d_model=128, n_layer=2)
    sd = random_init_state_dict(cfg, seed=11)
    rng = np.random.default_rng(3)
    L = 96
    window = "".join(rng.choice(list("ACGT"), size=L))
    window = window[:10] + "N" + window[11:40] + "a" + window[41:]   # an N (not mutated) and a lower-case base
    model = CaduceusForMaskedLM.from_pretrained(sd, config=cfg, torch_dtype=torch.float32).to(cuda_device)
    res = saturation_mutagenesis(model, window, batch_size=40)
    assert len(res["pos"]) == 3 * (L - 1) and 10 not in set(res["pos"].tolist())
    # oracle: one masked forward per position
    tok = model._tokenizer
    base_ids = torch.from_numpy(tok.encode_bytes(np.frombuffer(window.encode(), dtype=np.uint8)).astype(np.int64))
    want = {}
    for p in range(L):
        if p == 10:
            continue
        ids = base_ids.clone()[None, :]
        ids[0, p] = tok.mask_token_id
        lg, _ = O.caduceus_forward(sd, cfg, ids, dtype=torch.float32)
        want[p] = lg[0, p, 3:7].numpy()
    scale = max(np.abs(v).max() for v in want.values())
    for pos, ref, alt, score in zip(res["pos"], res["ref"], res["alt"], res["score"]):
        w = want[int(pos)]
        expect = w["ACGT".index(chr(alt))] - w["ACGT".index(chr(ref))]
        assert abs(score - expect) <= 2e-4 * scale + 1e-5


@pytest.mark.parametrize("L", [8192, 1000])
def test_long_context_hidden_states_match_oracle(cuda_device, L):
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    cfg = CaduceusConfig(d_model=128, n_layer=2)
    sd = random_init_state_dict(cfg, seed=12)
    g = torch.Generator().manual_seed(L)
    ids = torch.randint(3, 7, (1, L), generator=g)
    ids[0, L // 2] = 1
    want_logits, want_hs = O.caduceus_forward(sd, cfg, ids, dtype=torch.float32, output_hidden_states=True)
    model = CaduceusForMaskedLM.from_pretrained(sd, config=cfg, torch_dtype=torch.float32).to(cuda_device)
    out = model(input_ids=ids.to(cuda_device), output_hidden_states=True)
    hs = out.hidden_states[-1].cpu()
    assert hs.shape == (1, L, 2 * cfg.d_model)
    assert ((hs - want_hs[-1]).abs().max() / want_hs[-1].abs().max()).item() <= 1e-4
    assert ((out.logits.cpu() - want_logits).abs().max() / want_logits.abs().max()).item() <= 1e-4
    # embedding tap of the reference (train_XGBoost.py:104-113): average the forward half and the channel-reversed RC half
    emb = hs[:, L // 2, :]
    d = cfg.d_model
    avg = (emb[:, :d] + emb[:, d:].flip(-1)) / 2
    want_emb = want_hs[-1][:, L // 2, :]
    want_avg = (want_emb[:, :d] + want_emb[:, d:].flip(-1)) / 2
    assert torch.allclose(avg, want_avg, rtol=1e-3, atol=1e-4)
