"""More pins of the oracle (and of libpcad, directly) to the reference's own kernel families, beside
tests/test_thirdparty_pin_gpu.py.  The reference's arithmetic lives in ``mamba-ssm``, absent here; this
image carries three pieces of code that descend from it, used as CHECKERS ONLY (never on the product path):

* ``flash_attn.ops.triton.layer_norm.rms_norm_fn`` -- the Triton fused add + RMSNorm that ``mamba_ssm/ops/triton/
  layer_norm.py`` is a copy of (same author, same file): the exact call the Caduceus block makes with
  ``fused_add_norm=True`` (``prenorm=True``, ``residual_in_fp32``) -> oracle.rms_norm_add, pcad_op_add_rmsnorm;
* vLLM's ``mamba_chunk_scan_combined_varlen`` -- port of ``mamba_ssm/ops/triton/ssd_combined.py``, the Mamba-2 chunked
  SSD scan (``dt_softplus``, ``dt_bias``, ``D``; ``z=None`` because Mamba2 gates in its RMSNormGated) ->
  oracle.ssd_scan_ref, pcad_op_ssd_scan;
* vLLM's ``rms_norm_gated`` -- port of ``mamba_ssm/ops/triton/layernorm_gated.py`` (``norm_before_gate=False``) ->
  oracle.rmsnorm_gated, pcad_op_gated_norm_sum.

A third-party call that cannot be made on this build (import error, refused arguments) SKIPS with the reason; a numeric
disagreement FAILS.
"""
import ctypes as C

import pytest
import torch

from oracle import caduceus_oracle as O

pytestmark = pytest.mark.gpu

BF16, F32 = 0, 1


def ptr(t):
    return C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


@pytest.fixture(scope="module")
def lib(cuda_device):
    from plantcaduceus_b200 import _lib
    return _lib.load()


# ---- fused add + RMSNorm: the Triton kernel mamba_ssm's block calls ------------------------------------------------------
@pytest.mark.parametrize("rows,d,td,res32", [(300, 384, torch.float32, True), (64, 1024, torch.bfloat16, False),
                                             (130, 512, torch.bfloat16, True)])
def test_add_rmsnorm_matches_rms_norm_fn(lib, cuda_device, rows, d, td, res32):
    try:
        from flash_attn.ops.triton.layer_norm import rms_norm_fn
    except Exception as e:  # pragma: no cover - depends on the image
        pytest.skip(f"flash_attn rms_norm_fn not importable: {type(e).__name__}: {e}")
    g = torch.Generator().manual_seed(rows + d)
    rd = torch.float32 if res32 else td
    x = torch.randn(rows, d, generator=g).to(td)
    res = (2 * torch.randn(rows, d, generator=g)).to(rd)
    w = (1 + 0.5 * torch.randn(d, generator=g)).to(td).float()     # representable in the model dtype
    xd, rdv, wd = x.to(cuda_device), res.to(cuda_device), w.to(cuda_device)
    try:
        y3, r3 = rms_norm_fn(xd, wd.to(td), None, residual=rdv, eps=1e-5, prenorm=True, residual_in_fp32=res32)
        torch.cuda.synchronize()
    except Exception as e:
        pytest.skip(f"rms_norm_fn refused the call: {type(e).__name__}: {str(e)[:200]}")
    assert y3.dtype == td and r3.dtype == rd
    oy, ores = O.rms_norm_add(x, res, w.to(td), 1e-5, res32)
    assert ores.dtype == rd and oy.dtype == td
    ulp = 2 ** -7 if td == torch.bfloat16 else 1e-6
    scale = oy.float().abs().max().item()
    assert (r3.cpu().float() - ores.float()).abs().max().item() <= (ulp if rd == torch.bfloat16 else 1e-6) * ores.float().abs().max().item()
    assert (y3.cpu().float() - oy.float()).abs().max().item() <= 2 * ulp * scale
    # the engine's kernel against the same call
    y = torch.empty_like(xd)
    r_out = torch.empty_like(rdv)
    code = lambda t: F32 if t == torch.float32 else BF16
    rc = lib.pcad_op_add_rmsnorm(ptr(xd), ptr(rdv), ptr(wd), ptr(y), ptr(r_out), rows, d, C.c_float(1e-5), code(td), code(rd),
                                 stream())
    assert rc == 0, lib.pcad_last_error(None)
    torch.cuda.synchronize()
    assert (r_out.float() - r3.float()).abs().max().item() <= (ulp if rd == torch.bfloat16 else 1e-6) * r3.float().abs().max().item()
    assert (y.float() - y3.float()).abs().max().item() <= 2 * ulp * scale


# ---- Mamba-2 SSD scan: port of mamba_ssm's mamba_chunk_scan_combined -----------------------------------------------------
def _vllm_ssd(x, dt, A, Bm, Cm, D, dt_bias, chunk):
    """x [L, H, P], dt [L, H], Bm / Cm [L, 1, N], fp32 CUDA, ONE sequence.  Returns y [L, H, P]."""
    try:
        from vllm.model_executor.layers.mamba.ops.ssd_combined import mamba_chunk_scan_combined_varlen as fn
    except Exception as e:  # pragma: no cover
        pytest.skip(f"vLLM mamba_chunk_scan_combined_varlen not importable: {type(e).__name__}: {e}")
    L = x.shape[0]
    dev = x.device
    bounds = list(range(0, L, chunk)) + [L]
    cu_chunks = torch.tensor(bounds, dtype=torch.int32, device=dev)
    n_chunks = len(bounds) - 1
    out = torch.full_like(x, float("nan"))
    try:
        fn(x, dt, A, Bm, Cm, chunk, cu_seqlens=torch.tensor([0, L], dtype=torch.int32, device=dev),
           cu_chunk_seqlens=cu_chunks, last_chunk_indices=torch.tensor([n_chunks - 1], dtype=torch.int32, device=dev),
           seq_idx=torch.zeros(n_chunks, dtype=torch.int32, device=dev), out=out, D=D, z=None, dt_bias=dt_bias,
           dt_softplus=True, state_dtype=torch.float32)
        torch.cuda.synchronize()
    except Exception as e:
        pytest.skip(f"mamba_chunk_scan_combined_varlen refused the call: {type(e).__name__}: {str(e)[:300]}")
    return out


@pytest.mark.parametrize("S,L,H,chunk", [(2, 256, 4, 64), (1, 512, 2, 128), (1, 200, 2, 64)])
def test_ssd_scan_matches_mamba_ssm_chunk_scan_port(lib, cuda_device, S, L, H, chunk):
    """oracle.ssd_scan_ref and pcad_op_ssd_scan (fp32 sequential kernel) against the mamba_ssm-derived chunked scan.
    The Triton kernels multiply through tl.dot (TF32 for fp32 operands), so the bar is 1e-2 of the output scale: any
    difference of SEMANTICS (where dt_bias / softplus / D / the decay enter) is O(1)."""
    P, N = 64, 64
    E = H * P
    g = torch.Generator().manual_seed(S * 1000 + L + H)
    mk = lambda *s: torch.randn(*s, generator=g)
    x, Bm, Cm = mk(S, L, H, P), mk(S, L, 1, N) / N ** 0.5, mk(S, L, 1, N)
    dt = mk(S, L, H) * 0.5
    A = -(torch.rand(H, generator=g) * 4 + 0.5)
    D = 1 + 0.1 * mk(H)
    bias = mk(H) - 2
    want = O.ssd_scan_ref(x, dt, A, Bm, Cm, D, bias)                                   # [S, L, H, P]
    dev = lambda t: t.to(cuda_device).contiguous()
    got = torch.stack([_vllm_ssd(dev(x[s]), dev(dt[s]), dev(A), dev(Bm[s]), dev(Cm[s]), dev(D), dev(bias), chunk)
                       for s in range(S)]).cpu()
    assert not torch.isnan(got).any()
    scale = want.abs().max().item()
    assert (got - want).abs().max().item() <= 1e-2 * scale
    # libpcad's fp32 kernel, forward direction, against the same port (no oracle in between)
    CD = E + 2 * N
    xbc = torch.cat([x.reshape(S * L, E), Bm.reshape(S * L, N), Cm.reshape(S * L, N)], dim=1)
    assert xbc.shape[1] == CD
    xd, dtd = dev(xbc), dev(dt.reshape(S * L, H))
    Ad, Dd, bd = dev(A), dev(D), dev(bias)
    y = [torch.full((S * L, E), float("nan"), device=cuda_device) for _ in range(2)]
    rc = lib.pcad_op_ssd_scan(ptr(xd), ptr(xd), CD, ptr(dtd), H, ptr(Ad), ptr(Dd), ptr(bd), ptr(Ad), ptr(Dd), ptr(bd),
                              ptr(y[0]), ptr(y[1]), S, L, H, F32, 1, stream())
    assert rc == 0, lib.pcad_last_error(None)
    torch.cuda.synchronize()
    assert (y[0].cpu().reshape(S, L, H, P) - got).abs().max().item() <= 1e-2 * scale
    # the reverse direction is the same scan on the flipped sequence
    got_r = torch.stack([_vllm_ssd(dev(x[s].flip(0)), dev(dt[s].flip(0)), dev(A), dev(Bm[s].flip(0)), dev(Cm[s].flip(0)),
                                   dev(D), dev(bias), chunk).flip(0) for s in range(S)]).cpu()
    assert (y[1].cpu().reshape(S, L, H, P) - got_r).abs().max().item() <= 1e-2 * scale


# ---- Mamba-2 gated RMSNorm: port of mamba_ssm's layernorm_gated ----------------------------------------------------------
@pytest.mark.parametrize("rows,E", [(200, 1536), (33, 2048)])
def test_gated_norm_matches_mamba_ssm_layernorm_gated_port(lib, cuda_device, rows, E):
    try:
        from vllm.model_executor.layers.mamba.ops.layernorm_gated import rms_norm_gated
    except Exception as e:  # pragma: no cover
        pytest.skip(f"vLLM rms_norm_gated not importable: {type(e).__name__}: {e}")
    g = torch.Generator().manual_seed(rows + E)
    yf, yr, z = (torch.randn(rows, E, generator=g) for _ in range(3))
    wf, wr = 1 + 0.3 * torch.randn(E, generator=g), 1 + 0.3 * torch.randn(E, generator=g)
    dev = lambda t: t.to(cuda_device).contiguous()
    a, b, zz, w1, w2 = dev(yf), dev(yr), dev(z), dev(wf), dev(wr)
    try:
        n_f = rms_norm_gated(a, w1, None, z=zz, eps=1e-5, group_size=None, norm_before_gate=False)
        n_r = rms_norm_gated(b, w2, None, z=zz, eps=1e-5, group_size=None, norm_before_gate=False)
        torch.cuda.synchronize()
    except Exception as e:
        pytest.skip(f"rms_norm_gated refused the call: {type(e).__name__}: {str(e)[:200]}")
    want_f = O.rmsnorm_gated(yf, z, wf, 1e-5)
    scale = want_f.abs().max().item()
    assert (n_f.cpu() - want_f).abs().max().item() <= 1e-5 * scale
    out = torch.full((rows, E), float("nan"), device=cuda_device)
    rc = lib.pcad_op_gated_norm_sum(ptr(a), ptr(b), ptr(zz), E, ptr(w1), ptr(w2), ptr(out), rows, E, C.c_float(1e-5), F32, stream())
    assert rc == 0, lib.pcad_last_error(None)
    torch.cuda.synchronize()
    assert (out - (n_f + n_r)).abs().max().item() <= 2e-5 * scale
