"""The one numeric known-answer the reference holds for this path: the stored output of its
``notebooks/examples.ipynb`` -- softmax over the a,c,g,t logits at the masked index 255 of the README sequence, from the
PRETRAINED ``kuleshov-group/PlantCaduceus_l20`` checkpoint: ``[0.96960527, 0.00782286, 0.01123959, 0.01133224]``
(tests/golden/notebook_example.json holds the sequence and the numbers).

It needs the hub checkpoint, which is not reachable offline: these tests are OPT-IN.  Point ``PCAD_PLANTCADUCEUS_L20_DIR`` at a
local snapshot of that repository (config.json, model.safetensors or pytorch_model.bin, tokenizer.json) and they pin

* the CPU oracle (``oracle.caduceus_forward``: every [EXT] assumption of SURVEY.md Appendix A at once -- half ordering, tied
  head, token ids, residual dtype, epsilon, delta_bias / softplus placement) -- runs without a GPU;
* the engine through ``CaduceusForMaskedLM.from_pretrained(dir)`` in fp32 and bf16 (``-m gpu``).

Without the variable they skip, and the oracle stays "parity unpinned" in the prompt's sense.
"""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "notebook_example.json")
CKPT = os.environ.get("PCAD_PLANTCADUCEUS_L20_DIR")

needs_ckpt = pytest.mark.skipif(not (CKPT and os.path.isdir(CKPT)),
                                reason="set PCAD_PLANTCADUCEUS_L20_DIR to a local snapshot of kuleshov-group/PlantCaduceus_l20")


def _case():
    with open(GOLD) as f:
        g = json.load(f)
    return g["sequence"], int(g["pos"]), np.array(g["probs"], dtype=np.float32)


def test_notebook_fixture_is_well_formed():
    seq, pos, probs = _case()
    assert len(seq) == 512 and set(seq) <= set("ACGT") and seq[pos] == "A"
    assert abs(float(probs.sum()) - 1.0) < 1e-6 and int(probs.argmax()) == "acgt".index(seq[pos].lower())


@needs_ckpt
def test_oracle_reproduces_notebook_probabilities():
    from oracle import caduceus_oracle as O
    from plantcaduceus_b200 import CharDNATokenizer
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    seq, pos, want = _case()
    m = CaduceusForMaskedLM.from_pretrained(CKPT, torch_dtype=torch.float32)      # parses config + weights; stays on the CPU
    tok = CharDNATokenizer.from_pretrained(CKPT)
    ids = tok.encode_plus(seq, return_tensors="pt")["input_ids"]
    assert ids.shape == (1, 512)
    ids[0, pos] = tok.mask_token_id
    sd = {k: v.float() for k, v in m.state_dict().items() if v.is_floating_point()}
    with torch.inference_mode():
        logits, hs = O.caduceus_forward(sd, m.config, ids, dtype=torch.float32, output_hidden_states=True)
    assert tuple(hs[-1].shape) == (1, 512, 2 * m.config.d_model) == (1, 512, 768)
    v = tok.get_vocab()
    got = O.extract_acgt_probs(logits, pos, [v[c] for c in "acgt"]).numpy()[0]
    assert np.abs(got - want).max() <= 1e-3, (got, want)


@pytest.mark.gpu
@needs_ckpt
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 2e-2)])
def test_engine_reproduces_notebook_probabilities(cuda_device, dtype, tol):
    from plantcaduceus_b200 import CharDNATokenizer
    from plantcaduceus_b200 import genome_io as gio
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    seq, pos, want = _case()
    tok = CharDNATokenizer.from_pretrained(CKPT)
    m = CaduceusForMaskedLM.from_pretrained(CKPT, torch_dtype=dtype)
    m.set_tokenizer(tok)
    m.to(cuda_device)
    ascii_row = torch.from_numpy(tok.windows_to_ascii([seq], 512))
    got = gio.softmax4(m.score_windows_host(ascii_row, pos).numpy())[0]
    assert np.abs(got - want).max() <= tol, (got, want)
