"""CPU oracle for the PlantCaduceus (Caduceus / RC-equivariant BiMamba; Mamba-1, and Mamba-2 for PlantCAD2) forward pass.

TEST INFRASTRUCTURE ONLY.  Nothing under ``plantcaduceus_b200/`` may import this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` as the checker / CPU baseline, never as the product path.

PARITY UNPINNED: the arithmetic of this path lives in dependencies that are not vendored in
/root/reference and cannot be imported in the build container:
  * HF-hub remote code ``kuleshov-group/PlantCaduceus_l{20,24,28,32}`` (modeling_caduceus.py,
    modeling_rcps.py), loaded at reference src/zero_shot_score.py:91;
  * ``mamba-ssm==2.2.2`` and ``causal-conv1d==1.4.0`` (reference env/requirements.txt:9-10).
The reference ships no tests or golden vectors for this path.  This file
restates the published algorithm of those packages (Mamba.forward slow path with
``selective_scan_ref``; Caduceus RCPS wrappers) with *materialised* flips / cats so it is
structurally independent of the engine's index math.  It is pinned by:
  (1) reverse-complement equivariance (holds for any weights; tests/test_oracle.py),
  (2) an independent implementation of the Mamba-1 mixer that IS importable here,
      ``transformers.models.mamba.modeling_mamba.MambaMixer.slow_forward`` (tests/test_oracle.py),
  (3) parameter counts against reference README.md:60-63,
  (4) the module tree / shapes printed at reference notebooks/examples.ipynb:61-98,132,183,
  (5) on the GPU box, code that descends from the reference's own kernels (vLLM's ports of mamba_ssm's selective_scan_fwd,
      mamba_chunk_scan_combined and layernorm_gated; flash_attn's rms_norm_fn, the file mamba_ssm's layer_norm.py copies):
      tests/test_thirdparty_pin_gpu.py, tests/test_thirdparty_pin2_gpu.py,
  (6) opt-in, the reference notebook's stored softmax output of the pretrained l20 checkpoint
      (tests/test_reference_notebook_pin.py; needs a local snapshot of the hub checkpoint).
What IS pinned by outputs of the reference itself: everything of the path that lives in /root/reference -- the scoring
functions at the bottom of this file and the host path / callers they restate -- through tests/golden/reference_run/
(the reference's zero_shot_score.py main(), seq_from_vcf, zero-shot-eval.py, train_XGBoost.extract_embeddings EXECUTED in
the build container with this oracle as the model behind them; tests/golden/make_reference_run_golden.py).

Every function cites what it follows.  "[EXT]" marks upstream code that is not in /root/reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# Mamba-1 mixer  [EXT mamba_ssm/modules/mamba_simple.py: Mamba.forward, non-fused branch;
#                 mamba_ssm/ops/selective_scan_interface.py: selective_scan_ref]
# --------------------------------------------------------------------------------------------
def selective_scan_ref(u, delta, A, B, C, D, z, delta_bias):
    """u, delta, z: [b, E, L]; A: [E, N] fp32; B, C: [b, N, L]; D, delta_bias: [E] fp32.

    fp32 state, softplus on (delta + bias), D skip, SiLU(z) gate, output cast to u.dtype.
    """
    dtype_in = u.dtype
    u = u.float()
    delta = delta.float() + delta_bias[None, :, None].float()
    delta = F.softplus(delta)  # threshold 20 -> identity above, same as the CUDA kernel
    B = B.float()
    C = C.float()
    b, E, L = u.shape
    N = A.shape[1]
    x = torch.zeros((b, E, N), dtype=torch.float32)
    ys = []
    for i in range(L):
        dA = torch.exp(delta[:, :, i, None] * A[None])                       # [b, E, N]
        dBu = (delta[:, :, i] * u[:, :, i])[:, :, None] * B[:, None, :, i]   # [b, E, N]
        x = dA * x + dBu
        ys.append((x * C[:, None, :, i]).sum(-1))
    y = torch.stack(ys, dim=2)
    y = y + u * D[None, :, None].float()
    y = y * F.silu(z.float())
    return y.to(dtype_in)


def mamba_mixer(u, p: Dict[str, torch.Tensor]):
    """One direction of Mamba-1 on u [b, L, d] (dtype = model dtype). p holds this direction's tensors."""
    b, L, d = u.shape
    E = p["conv1d.weight"].shape[0]
    N = p["A_log"].shape[1]
    R = p["dt_proj.weight"].shape[1]
    xz = F.linear(u, p["in_proj.weight"]).transpose(1, 2)                    # [b, 2E, L]
    x, z = xz.chunk(2, dim=1)
    A = -torch.exp(p["A_log"].float())
    # depthwise causal conv, left zero padding K-1, truncated to L, then SiLU
    # (causal_conv1d accumulates in fp32 and rounds once to the I/O dtype)
    xc = F.conv1d(x.float(), p["conv1d.weight"].float(), p["conv1d.bias"].float(),
                  padding=p["conv1d.weight"].shape[-1] - 1, groups=E)[..., :L]
    x = F.silu(xc).to(u.dtype)
    x_dbl = F.linear(x.transpose(1, 2).reshape(b * L, E), p["x_proj.weight"])  # [(b L), R+2N]
    dt, Bm, Cm = torch.split(x_dbl, [R, N, N], dim=-1)
    dt = (p["dt_proj.weight"] @ dt.t()).reshape(E, b, L).transpose(0, 1)       # [b, E, L], bias NOT added here
    Bm = Bm.reshape(b, L, N).transpose(1, 2).contiguous()
    Cm = Cm.reshape(b, L, N).transpose(1, 2).contiguous()
    y = selective_scan_ref(x, dt, A, Bm, Cm, p["D"].float(), z, p["dt_proj.bias"].float())
    return F.linear(y.transpose(1, 2), p["out_proj.weight"])


# --------------------------------------------------------------------------------------------
# Mamba-2 / SSD mixer  [EXT mamba_ssm/modules/mamba2.py: Mamba2.forward, non-fused branch
#                       (use_mem_eff_path=False); mamba_ssm/ops/triton/ssd_combined.py:
#                       mamba_chunk_scan_combined == the recurrence below (ssd_minimal / selective_state_update);
#                       mamba_ssm/ops/triton/layernorm_gated.py: rmsnorm_fn(norm_before_gate=False)]
# PlantCAD2 checkpoints (reference docs/PlantCAD2-overview.md:17-21, src/zero-shot-eval.py:54-72) use this mixer.
# --------------------------------------------------------------------------------------------
def ssd_scan_ref(x, dt, A, B, C, D, dt_bias):
    """x: [b, L, H, P]; dt: [b, L, H] (raw); A, D, dt_bias: [H] fp32; B, C: [b, L, G, N].  Sequential state-space
    recurrence with scalar decay per head: dt = softplus(dt + dt_bias); S <- exp(dt A) S + dt x (x) B; y = S C + D x.
    fp32 state, output cast to x.dtype (the SSD kernels accumulate in fp32 and store the I/O dtype)."""
    dtype_in = x.dtype
    b, L, H, P = x.shape
    G, N = B.shape[2], B.shape[3]
    x = x.float()
    dt = F.softplus(dt.float() + dt_bias.float()[None, None, :])
    Bh = B.float().repeat_interleave(H // G, dim=2)          # [b, L, H, N]
    Ch = C.float().repeat_interleave(H // G, dim=2)
    S = torch.zeros((b, H, P, N), dtype=torch.float32)
    ys = []
    for t in range(L):
        dA = torch.exp(dt[:, t] * A[None, :])                                   # [b, H]
        dBx = (dt[:, t, :, None] * x[:, t])[..., None] * Bh[:, t, :, None, :]     # [b, H, P, N]
        S = dA[:, :, None, None] * S + dBx
        ys.append((S * Ch[:, t, :, None, :]).sum(-1))                           # [b, H, P]
    y = torch.stack(ys, dim=1) + x * D.float()[None, None, :, None]
    return y.to(dtype_in)


def rmsnorm_gated(y, z, weight, eps):
    """[EXT RMSNormGated(norm_before_gate=False, group_size=d_inner)]: fp32 y * silu(z), RMS over the whole row, * weight."""
    xf = y.float() * F.silu(z.float())
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (xf * rstd * weight.float()).to(y.dtype)


def mamba2_mixer(u, p: Dict[str, torch.Tensor], headdim: int = 64, ngroups: int = 1, eps: float = 1e-5):
    """One direction of Mamba-2 on u [b, L, d].  in_proj rows are ordered z | x | B | C | dt."""
    b, L, d = u.shape
    E = p["norm.weight"].shape[0]
    H = p["A_log"].shape[0]
    CD = p["conv1d.weight"].shape[0]
    N = (CD - E) // (2 * ngroups)
    zxbcdt = F.linear(u, p["in_proj.weight"])                                  # [b, L, 2E + 2GN + H]
    z, xBC, dt = torch.split(zxbcdt, [E, CD, H], dim=-1)
    xc = F.conv1d(xBC.transpose(1, 2).float(), p["conv1d.weight"].float(), p["conv1d.bias"].float(),
                  padding=p["conv1d.weight"].shape[-1] - 1, groups=CD)[..., :L]
    xBC = F.silu(xc).transpose(1, 2).to(u.dtype)
    x, Bm, Cm = torch.split(xBC, [E, ngroups * N, ngroups * N], dim=-1)
    A = -torch.exp(p["A_log"].float())
    y = ssd_scan_ref(x.reshape(b, L, H, headdim), dt, A, Bm.reshape(b, L, ngroups, N), Cm.reshape(b, L, ngroups, N),
                     p["D"].float(), p["dt_bias"].float())
    y = rmsnorm_gated(y.reshape(b, L, E), z, p["norm.weight"], eps)
    return F.linear(y, p["out_proj.weight"])


# --------------------------------------------------------------------------------------------
# Caduceus wrappers  [EXT HF-hub modeling_caduceus.py / modeling_rcps.py]
# --------------------------------------------------------------------------------------------
def bimamba(u, p_fwd, p_rev, mixer=None):
    """[EXT BiMambaWrapper.forward], strategy "add": fwd(u) + flip_L(rev(flip_L(u)))."""
    mixer = mixer or mamba_mixer
    out = mixer(u, p_fwd)
    out_rev = mixer(u.flip(dims=(1,)), p_rev).flip(dims=(1,))
    return out + out_rev


def rcps_wrapper(x, p_fwd, p_rev, mixer=None):
    """[EXT RCPSWrapper.forward]: same submodule on the fwd half and on flip_{L,D}(RC half)."""
    d = x.shape[-1] // 2
    fwd_out = bimamba(x[..., :d], p_fwd, p_rev, mixer)
    rc_out = bimamba(torch.flip(x[..., d:], dims=[-2, -1]), p_fwd, p_rev, mixer)
    return torch.cat([fwd_out, torch.flip(rc_out, dims=[-2, -1])], dim=-1)


def rms_norm_add(x, residual, weight, eps, residual_in_fp32, prenorm=True):
    """[EXT mamba_ssm/ops/triton/layer_norm.py rms_norm_fn]: fp32 (x + residual), residual_out stored in
    fp32 iff residual_in_fp32 else x.dtype, stats from the un-rounded fp32 sum, y in x.dtype."""
    xf = x.float()
    if residual is not None:
        xf = xf + residual.float()
    res_dtype = torch.float32 if residual_in_fp32 else x.dtype
    rstd = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    y = (xf * rstd * weight.float()).to(x.dtype)
    return (y, xf.to(res_dtype)) if prenorm else y


def rcps_add_norm(hidden, residual, weight, eps, residual_in_fp32, prenorm=True):
    """[EXT RCPSMambaBlock.forward, fused_add_norm branch]: norm the fwd half, and the RC half flipped."""
    d = hidden.shape[-1] // 2
    r_f = None if residual is None else residual[..., :d]
    r_r = None if residual is None else residual[..., d:].flip(dims=[-2, -1])
    out_f = rms_norm_add(hidden[..., :d], r_f, weight, eps, residual_in_fp32, prenorm)
    out_r = rms_norm_add(hidden[..., d:].flip(dims=[-2, -1]), r_r, weight, eps, residual_in_fp32, prenorm)
    if prenorm:
        h = torch.cat([out_f[0], out_r[0].flip(dims=[-2, -1])], dim=-1)
        r = torch.cat([out_f[1], out_r[1].flip(dims=[-2, -1])], dim=-1)
        return h, r
    return torch.cat([out_f, out_r.flip(dims=[-2, -1])], dim=-1)


def rcps_embedding(input_ids, emb_weight, comp):
    """[EXT RCPSEmbedding.forward]: cat[E[ids], flip_{L,D}(E[comp[flip_L(ids)]])]."""
    fwd = F.embedding(input_ids, emb_weight)
    rc_ids = comp[torch.flip(input_ids, dims=[-1])]
    rc = torch.flip(F.embedding(rc_ids, emb_weight), dims=[-2, -1])
    return torch.cat([fwd, rc], dim=-1)


def rcps_lm_head(x, weight, comp):
    """[EXT RCPSLMHead.forward]: x[..,:d] W^T + flip_D(x[..,d:]) W[comp]^T."""
    d = x.shape[-1] // 2
    fwd = F.linear(x[..., :d], weight)
    rc = F.linear(torch.flip(x[..., d:], dims=[-1]), weight[comp, :])
    return fwd + rc


_M1_NAMES = ("in_proj.weight", "conv1d.weight", "conv1d.bias", "x_proj.weight", "dt_proj.weight",
             "dt_proj.bias", "A_log", "D", "out_proj.weight")
_M2_NAMES = ("in_proj.weight", "conv1d.weight", "conv1d.bias", "dt_bias", "A_log", "D", "norm.weight", "out_proj.weight")


def _dir_params(sd, i, direction, dtype, names=_M1_NAMES):
    pre = f"caduceus.backbone.layers.{i}.mixer.submodule.{direction}."
    out = {}
    for name in names:
        t = sd[pre + name]
        # from_pretrained(torch_dtype=dtype) casts every parameter; Mamba reads A_log / D (/ dt_bias) back with .float()
        out[name] = t.to(dtype).float() if name in ("A_log", "D", "dt_bias") else t.to(dtype)
    return out


def caduceus_forward(sd: Dict[str, torch.Tensor], cfg, input_ids: torch.Tensor,
                     dtype: torch.dtype = torch.float32,
                     output_hidden_states: bool = False) -> Tuple[torch.Tensor, Optional[List[torch.Tensor]]]:
    """[EXT CaduceusForMaskedLM.forward]: returns (logits fp32 [B, L, V], hidden_states list or None).

    ``hidden_states[-1]`` is the final normed [B, L, 2d] tensor (fwd half, then RC half with sequence
    and channels reversed back), which is what reference src/train_XGBoost.py:104-113 taps.
    """
    comp = torch.tensor([cfg.complement_map[i] for i in range(cfg.vocab_size)], dtype=torch.long)
    emb = sd["caduceus.backbone.embeddings.word_embeddings.embedding.weight"].to(dtype)
    head = sd["lm_head.lm_head.weight"].to(dtype)
    eps = cfg.norm_epsilon
    hidden = rcps_embedding(input_ids, emb, comp)
    residual = None
    mamba2 = getattr(cfg, "is_mamba2", False)
    names = _M2_NAMES if mamba2 else _M1_NAMES
    mixer = (lambda u, p: mamba2_mixer(u, p, cfg.headdim, cfg.ngroups)) if mamba2 else mamba_mixer
    all_hidden = [] if output_hidden_states else None
    for i in range(cfg.n_layer):
        if output_hidden_states:
            all_hidden.append(hidden)
        w = sd[f"caduceus.backbone.layers.{i}.norm.weight"].to(dtype)
        hidden, residual = rcps_add_norm(hidden, residual, w, eps, cfg.residual_in_fp32, prenorm=True)
        hidden = rcps_wrapper(hidden, _dir_params(sd, i, "mamba_fwd", dtype, names), _dir_params(sd, i, "mamba_rev", dtype, names),
                              mixer)
    w_f = sd["caduceus.backbone.norm_f.weight"].to(dtype)
    hidden = rcps_add_norm(hidden, residual, w_f, eps, cfg.residual_in_fp32, prenorm=False)
    if output_hidden_states:
        all_hidden.append(hidden)
    logits = rcps_lm_head(hidden, head, comp).float()
    return logits, all_hidden


# --------------------------------------------------------------------------------------------
# Scoring  [reference src/zero_shot_score.py:107-134]
# --------------------------------------------------------------------------------------------
def extract_acgt_probs(logits: torch.Tensor, token_idx: int, acgt_ids) -> torch.Tensor:
    """softmax over the 4 nucleotide logits at the masked index (reference :117-119)."""
    sel = logits[:, token_idx, list(acgt_ids)]
    return torch.softmax(sel.float().cpu(), dim=1)


def zero_shot_llr(probs, refs, alts):
    """log(p_alt / p_ref) per row (reference :124-134); refs/alts are 'A','C','G','T' strings."""
    import numpy as np
    nuc = ["A", "C", "G", "T"]
    p = probs.numpy() if hasattr(probs, "numpy") else probs
    return [float(np.log(p[i][nuc.index(a)] / p[i][nuc.index(r)])) for i, (r, a) in enumerate(zip(refs, alts))]


def reverse_complement_ids(input_ids: torch.Tensor, cfg) -> torch.Tensor:
    comp = torch.tensor([cfg.complement_map[i] for i in range(cfg.vocab_size)], dtype=torch.long)
    return comp[torch.flip(input_ids, dims=[-1])]
