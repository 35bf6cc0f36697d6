"""CPU: the C-ABI library builds, loads and exports every symbol include/pfm_b200.h declares;
without a CUDA device the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "pfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    syms = header_symbols()
    for s in ("pfm_epic_create", "pfm_epic_set_weights", "pfm_epic_forward", "pfm_epic_sample", "pfm_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib_built):
    from particle_fm_b200 import _lib
    lib = _lib.load()
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/pfm_b200.h but not exported"
    assert set(header_symbols()) == set(_lib.exported_symbols())
    assert lib.pfm_version() == 100


def test_library_has_no_torch_dependency(lib_built):
    import subprocess
    out = subprocess.run(["ldd", lib_built], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out


def test_sm100a_sass_only(lib_built):
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "--list-elf", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(lib_built):
    from particle_fm_b200 import _lib
    lib = _lib.load()
    cfg = _lib.EpicCfgC(3, 3, 128, 10, 6, 32, 1, 1, 0, 0, 1e-2, 0.01)
    h = C.c_void_p()
    rc = lib.pfm_epic_create(C.byref(cfg), 0, C.byref(h))
    assert rc == -2 and not h.value                       # PFM_ERR_CUDA
    assert b"no CPU fallback" in lib.pfm_last_error()
    from particle_fm_b200.models.components.epic import EPiC_encoder
    net = EPiC_encoder(latent=4, input_dim=3, hid_d=8, feats=3, equiv_layers=1, frequencies=2, t_local_cat=True,
                       t_global_cat=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            net(torch.zeros(2, 5, 4), torch.zeros(2, 5, 3), None, torch.ones(2, 5, 1))
