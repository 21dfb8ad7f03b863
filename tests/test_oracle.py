"""Pins for the CPU oracle (oracle/caduceus_oracle.py) that do not need the reference's un-vendored
dependencies: RC equivariance, an independent Mamba-1 mixer, parameter counts, shapes."""
import math

import pytest
import torch

from oracle import caduceus_oracle as O
from plantcaduceus_b200 import CaduceusConfig, count_parameters, preset, random_init_state_dict

torch.set_num_threads(max(1, min(8, torch.get_num_threads())))


def small_cfg(**kw):
    base = dict(d_model=128, n_layer=2)
    base.update(kw)
    return CaduceusConfig(**base)


def rand_ids(B, L, seed=0, with_special=True):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(3, 7, (B, L), generator=g)
    if with_special:
        ids[:, L // 2] = 1          # [MASK]
        ids[0, 0] = 2               # [UNK] (an N in the window)
    return ids


@pytest.mark.parametrize("residual_in_fp32", [False, True])
def test_rc_equivariance_exact(residual_in_fp32):
    cfg = small_cfg(residual_in_fp32=residual_in_fp32)
    sd = random_init_state_dict(cfg, seed=3)
    ids = rand_ids(2, 48, seed=1)
    logits, hs = O.caduceus_forward(sd, cfg, ids, output_hidden_states=True)
    rc_ids = O.reverse_complement_ids(ids, cfg)
    logits_rc, hs_rc = O.caduceus_forward(sd, cfg, rc_ids, output_hidden_states=True)
    comp = torch.tensor([cfg.complement_map[i] for i in range(cfg.vocab_size)])
    # SURVEY.md 8c property (1)
    assert torch.allclose(logits_rc, logits.flip(1)[..., comp], atol=1e-5, rtol=1e-5)
    assert torch.allclose(hs_rc[-1], hs[-1].flip(1, 2), atol=1e-5, rtol=1e-5)
    assert logits.shape == (2, 48, 8) and logits.dtype == torch.float32
    assert hs[-1].shape == (2, 48, 2 * cfg.d_model)


def test_mixer_matches_transformers_mamba():
    """Independent second opinion: transformers' MambaMixer.slow_forward (same parameter names)."""
    from transformers import MambaConfig
    from transformers.models.mamba.modeling_mamba import MambaMixer
    cfg = small_cfg()
    sd = random_init_state_dict(cfg, seed=5)
    p = O._dir_params(sd, 0, "mamba_fwd", torch.float32)
    mc = MambaConfig(hidden_size=cfg.d_model, state_size=16, conv_kernel=4, expand=2, time_step_rank=cfg.dt_rank,
                     use_bias=False, use_conv_bias=True, hidden_act="silu", num_hidden_layers=1, vocab_size=8)
    mixer = MambaMixer(mc, layer_idx=0).eval()
    with torch.no_grad():
        mixer.in_proj.weight.copy_(p["in_proj.weight"])
        mixer.conv1d.weight.copy_(p["conv1d.weight"])
        mixer.conv1d.bias.copy_(p["conv1d.bias"])
        mixer.x_proj.weight.copy_(p["x_proj.weight"])
        mixer.dt_proj.weight.copy_(p["dt_proj.weight"])
        mixer.dt_proj.bias.copy_(p["dt_proj.bias"])
        mixer.A_log.copy_(p["A_log"])
        mixer.D.copy_(p["D"])
        mixer.out_proj.weight.copy_(p["out_proj.weight"])
        u = torch.randn(2, 40, cfg.d_model, generator=torch.Generator().manual_seed(0))
        want = mixer.slow_forward(u)
        got = O.mamba_mixer(u, p)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4), (got - want).abs().max()


def test_direction_shared_projections_identity():
    """SURVEY.md 8c property (3): Bi(u) == (y_f + flip(y_r)) W_out^T with one in_proj."""
    cfg = small_cfg()
    sd = random_init_state_dict(cfg, seed=7)
    pf = O._dir_params(sd, 1, "mamba_fwd", torch.float32)
    pr = O._dir_params(sd, 1, "mamba_rev", torch.float32)
    assert torch.equal(pf["in_proj.weight"], pr["in_proj.weight"])
    assert torch.equal(pf["out_proj.weight"], pr["out_proj.weight"])
    assert not torch.equal(pf["x_proj.weight"], pr["x_proj.weight"])


@pytest.mark.parametrize("name,millions", [("PlantCaduceus_l20", 20.87), ("PlantCaduceus_l24", 43.6),
                                           ("PlantCaduceus_l32", 225.36)])
def test_parameter_counts(name, millions):
    """reference README.md:60-63 (20M / 40M / 225M); exact figures SURVEY.md Appendix A."""
    cfg = preset(name)
    n = 0
    d, E, N, R, V = cfg.d_model, cfg.d_inner, cfg.d_state, cfg.dt_rank, cfg.vocab_size
    per_dir = E * 4 + E + (R + 2 * N) * E + E * R + E + E * N + E
    n = V * d + cfg.n_layer * (2 * E * d + d * E + 2 * per_dir + d) + d
    assert abs(n / 1e6 - millions) < 0.05, n
    if name == "PlantCaduceus_l20":
        assert count_parameters(random_init_state_dict(cfg, 0)) == n


def test_llr_equals_logit_difference():
    """SURVEY.md 8c property (2): log(p_alt/p_ref) == logit_alt - logit_ref."""
    cfg = small_cfg()
    sd = random_init_state_dict(cfg, seed=11)
    ids = rand_ids(4, 32, seed=2)
    logits, _ = O.caduceus_forward(sd, cfg, ids)
    probs = O.extract_acgt_probs(logits, 16, (3, 4, 5, 6))
    refs, alts = ["A", "C", "G", "T"], ["C", "G", "T", "A"]
    llr = O.zero_shot_llr(probs, refs, alts)
    for i, (r, a) in enumerate(zip(refs, alts)):
        diff = float(logits[i, 16, 3 + "ACGT".index(a)] - logits[i, 16, 3 + "ACGT".index(r)])
        assert math.isclose(llr[i], diff, rel_tol=1e-4, abs_tol=1e-5)


def test_bf16_mode_close_to_fp32():
    cfg = small_cfg()
    sd = random_init_state_dict(cfg, seed=13)
    ids = rand_ids(2, 64, seed=3)
    l32, _ = O.caduceus_forward(sd, cfg, ids, dtype=torch.float32)
    l16, _ = O.caduceus_forward(sd, cfg, ids, dtype=torch.bfloat16)
    assert l16.dtype == torch.float32
    # sanity only: the reference's own bf16 logits are bf16-quantised (ulp 0.0156 at |logit| in [2,4))
    assert (l32 - l16).abs().max() < 5e-2


# ---- Mamba-2 / SSD (PlantCAD2) -----------------------------------------------------------------------------------
def small_m2(**kw):
    base = dict(d_model=128, n_layer=2, ssm_cfg=dict(layer="Mamba2", d_state=64, d_conv=4, expand=2, headdim=64, ngroups=1,
                                                      conv_bias=True, bias=False))
    base.update(kw)
    return CaduceusConfig(**base)


def test_mamba2_mixer_matches_transformers_mamba2():
    """Independent second opinion on the Mamba-2 mixer: transformers' Mamba2Mixer.torch_forward (chunked SSD in plain
    torch; same parameter names, same z | x | B | C | dt projection order)."""
    from transformers import Mamba2Config
    from transformers.models.mamba2.modeling_mamba2 import Mamba2Mixer
    cfg = small_m2()
    sd = random_init_state_dict(cfg, seed=5)
    p = O._dir_params(sd, 0, "mamba_fwd", torch.float32, O._M2_NAMES)
    mc = Mamba2Config(hidden_size=cfg.d_model, state_size=64, conv_kernel=4, expand=2, head_dim=64, num_heads=cfg.nheads,
                      n_groups=1, use_bias=False, use_conv_bias=True, hidden_act="silu", num_hidden_layers=1, vocab_size=8,
                      chunk_size=32, layer_norm_epsilon=1e-5, rms_norm=True)
    mixer = Mamba2Mixer(mc, layer_idx=0).eval()
    with torch.no_grad():
        mixer.in_proj.weight.copy_(p["in_proj.weight"])
        mixer.conv1d.weight.copy_(p["conv1d.weight"])
        mixer.conv1d.bias.copy_(p["conv1d.bias"])
        mixer.dt_bias.copy_(p["dt_bias"])
        mixer.A_log.copy_(p["A_log"])
        mixer.D.copy_(p["D"])
        mixer.norm.weight.copy_(p["norm.weight"])
        mixer.out_proj.weight.copy_(p["out_proj.weight"])
        u = torch.randn(2, 80, cfg.d_model, generator=torch.Generator().manual_seed(0))    # 80 = 2.5 chunks of 32
        want = mixer.torch_forward(u)
        got = O.mamba2_mixer(u, p, headdim=64, ngroups=1)
    assert torch.allclose(got, want, atol=3e-5, rtol=1e-4), (got - want).abs().max()


def test_mamba2_rc_equivariance_and_shapes():
    cfg = small_m2()
    cfg.validate_supported()
    sd = random_init_state_dict(cfg, seed=3)
    ids = rand_ids(2, 40, seed=1)
    logits, hs = O.caduceus_forward(sd, cfg, ids, output_hidden_states=True)
    logits_rc, hs_rc = O.caduceus_forward(sd, cfg, O.reverse_complement_ids(ids, cfg), output_hidden_states=True)
    comp = torch.tensor([cfg.complement_map[i] for i in range(cfg.vocab_size)])
    assert torch.allclose(logits_rc, logits.flip(1)[..., comp], atol=1e-5, rtol=1e-5)
    assert torch.allclose(hs_rc[-1], hs[-1].flip(1, 2), atol=1e-5, rtol=1e-5)
    assert logits.shape == (2, 40, 8) and hs[-1].shape == (2, 40, 2 * cfg.d_model)


@pytest.mark.parametrize("name,millions", [("PlantCAD2-Small-l24-d0768", 88), ("PlantCAD2-Medium-l48-d1024", 311),
                                           ("PlantCAD2-Large-l48-d1536", 694)])
def test_plantcad2_parameter_counts(name, millions):
    """The Mamba-2 hyper-parameters are not in the reference repo; d_state 64 / headdim 64 / ngroups 1 / expand 2 with the
    in/out projections tied across directions reproduce the published sizes (reference img/PlantCAD2-difference.jpg:
    Small 88M, Medium 311M, Large 694M; layers / widths docs/PlantCAD2-overview.md:17-21) to the printed digit."""
    cfg = preset(name)
    cfg.validate_supported()
    d, E, H, CD, DIP = cfg.d_model, cfg.d_inner, cfg.nheads, cfg.conv_dim, cfg.d_in_proj
    per_dir = CD * 4 + CD + 3 * H + E
    n = 8 * d + cfg.n_layer * (DIP * d + d * E + 2 * per_dir + d) + d
    assert round(n / 1e6) == millions, n
    # the alternatives do not: d_state 128 (Mamba2's default) or untied projections
    n128 = n + cfg.n_layer * (2 * 64 * d + 2 * 2 * 64 * 5)
    assert round(n128 / 1e6) != millions
    if "Small" in name:
        assert count_parameters(random_init_state_dict(cfg, 0)) == n


# ---- the plain-C restatement of the byte / integer rules (oracle/host_rules.c) -----------------------------------------
def _host_rules():
    import ctypes as C
    import __graft_entry__ as entry
    lib = C.CDLL(entry.build_oracle())
    lib.pcad_oracle_extract_window.restype = C.c_int
    return lib, C


def test_c_oracle_windows_equal_reference_seq_from_vcf_and_host_rule():
    """oracle/host_rules.c::pcad_oracle_extract_window against (1) the windows the REFERENCE'S seq_from_vcf returned
    (tests/golden/reference_run/: example VCF, and a small genome with soft-masked / N bases at four tokenIdx values),
    (2) plantcaduceus_b200.genome_io.extract_window on random positions and chromosome lengths."""
    import json
    import os
    import numpy as np
    from plantcaduceus_b200 import genome_io as gio
    lib, C = _host_rules()
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

    def c_window(chrom: bytes, pos0: int, tidx: int, L: int = 512) -> bytes:
        out = (C.c_uint8 * L)()
        buf = (C.c_uint8 * max(len(chrom), 1)).from_buffer_copy(chrom or b"\0")
        lib.pcad_oracle_extract_window(buf, C.c_int64(len(chrom)), C.c_int64(pos0), tidx, L, out)
        return bytes(out)

    fasta = gio.read_fasta(os.path.join(gold, "reference_run", "small_genome.fa"))
    recs = gio.read_vcf(os.path.join(gold, "reference_run", "small.vcf"))[1]
    with open(os.path.join(gold, "reference_run", "small_windows.json")) as f:
        small = json.load(f)
    for tidx, want in small.items():
        got = [c_window(fasta[recs[i].chrom], recs[i].pos - 1, int(tidx)).decode() for i in want["record_indices"]]
        assert got == want["windows"], f"tokenIdx {tidx}"
    g = np.load(os.path.join(gold, "reference_run", "vcf_windows.npz"))
    fasta = gio.read_fasta(os.path.join(gold, "example_genome.fa.gz"))
    recs = gio.read_vcf(os.path.join(gold, "example_maize_snp.vcf"))[1]
    for k in range(0, len(g["record_indices"]), 7):
        r = recs[int(g["record_indices"][k])]
        assert c_window(fasta[r.chrom], r.pos - 1, 255) == bytes(g["windows"][k])
    rng = np.random.default_rng(4)
    for _ in range(300):
        n = int(rng.integers(1, 1500))
        chrom = bytes(rng.choice(np.frombuffer(b"ACGTacgtN", dtype=np.uint8), size=n))
        pos0, tidx = int(rng.integers(0, n)), int(rng.integers(0, 512))
        assert c_window(chrom, pos0, tidx) == gio.extract_window(chrom, pos0, tidx, 512)


def test_c_oracle_tokenise_rc_and_slop():
    import os
    import numpy as np
    from plantcaduceus_b200 import CharDNATokenizer
    lib, C = _host_rules()
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    tok = CharDNATokenizer()
    rows = [ln.rstrip("\n").split("\t")[6] for ln in open(os.path.join(gold, "example_snp.tsv"))][1:]
    ascii_mat = np.frombuffer("".join(rows).encode(), dtype=np.uint8).copy()
    ids = np.zeros_like(ascii_mat)
    p8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
    lut = np.ascontiguousarray(tok.lut, dtype=np.uint8)
    lib.pcad_oracle_tokenize(p8(ascii_mat), C.c_int64(ascii_mat.size), p8(lut), 512, 255, tok.mask_token_id, p8(ids))
    assert np.array_equal(ids.reshape(-1, 512), np.load(os.path.join(gold, "example_ids.npz"))["ids"])
    # RC strand ids == the oracle's reverse_complement_ids
    cfg = preset("PlantCaduceus_l20")
    comp = np.array([cfg.complement_map[i] for i in range(cfg.vocab_size)], dtype=np.uint8)
    row = np.ascontiguousarray(ids[:512])
    out = np.zeros(512, dtype=np.uint8)
    lib.pcad_oracle_rc_ids(p8(row), 512, p8(comp), p8(out))
    want = O.reverse_complement_ids(torch.from_numpy(row.astype(np.int64))[None], cfg)[0].numpy()
    assert np.array_equal(out, want)
    # format_VCF.sh interval == the start / end columns of the reference's example table
    s, e = C.c_int64(), C.c_int64()
    for ln in list(open(os.path.join(gold, "example_snp.tsv")))[1:20]:
        f = ln.split("\t")
        lib.pcad_oracle_slop(C.c_int64(int(f[3])), C.c_int64(10 ** 9), 255, 256, C.byref(s), C.byref(e))
        assert (s.value, e.value) == (int(f[1]), int(f[2]))
    lib.pcad_oracle_slop(C.c_int64(10), C.c_int64(100), 255, 256, C.byref(s), C.byref(e))
    assert (s.value, e.value) == (0, 100)
