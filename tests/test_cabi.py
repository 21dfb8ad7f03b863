"""The C-ABI library builds, loads on a CPU-only host and exports every symbol include/pcad.h declares.
No compute calls (there is no GPU here): creating a handle must fail cleanly with PCAD_ERR_CUDA, not crash."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pcad.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcad_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from plantcaduceus_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    from plantcaduceus_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.EXPORTS) == syms, (set(syms) ^ set(_lib.EXPORTS))
    for s in syms:
        assert hasattr(lib, s), f"libpcad.so does not export {s}"
    assert lib.pcad_abi_version() == _lib.ABI_VERSION == 3


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from plantcaduceus_b200 import _lib
    cfg = _lib.PcadConfig(128, 1, 8, 16, 4, 2, 8, 1e-5, 0, 0)
    for i in range(16):
        cfg.complement_map[i] = i
    h = C.c_void_p()
    rc = lib.pcad_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -2 and not h.value                     # PCAD_ERR_CUDA
    assert b"no CPU fallback" in lib.pcad_last_error(None)
    # unsupported configuration is rejected before any device work
    cfg.d_state = 8
    assert lib.pcad_create(C.byref(cfg), 0, C.byref(h)) == -1
    cfg.d_state, cfg.mixer = 16, 1                      # Mamba-2 needs d_state 64 / headdim 64 / ngroups 1
    assert lib.pcad_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert b"Mamba-2" in lib.pcad_last_error(None)
    from plantcaduceus_b200.modeling import CaduceusForMaskedLM
    from plantcaduceus_b200 import CaduceusConfig
    m = CaduceusForMaskedLM.from_random(CaduceusConfig(d_model=128, n_layer=1))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(input_ids=torch.zeros(1, 8, dtype=torch.long))


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: no module of the package (nor the CUDA sources) may import, load or mention it, and the
    package has no CPU fallback to route through."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "plantcaduceus_b200", "**", "*.py"), recursive=True) + \
        glob.glob(os.path.join(root, "plantcaduceus_b200", "csrc", "*.cu*"))
    assert len(files) > 15
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/.](caduceus_oracle|host_rules|_build)|libhost_rules", re.M)
    for f in files:
        assert not pat.search(open(f, errors="ignore").read()), f
