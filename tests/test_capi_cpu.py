"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol
include/mvae_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mvae_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from multimodal_vae_public_b200 import _lib, build
    build.build(verbose=False)
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mvae_b200.h but not exported"
    assert sorted(_lib.EXPORTED) == names, "ctypes binding and header disagree"
    assert lib.mvae_version() >= 100


def test_bad_arguments_return_error_codes_not_crash():
    from multimodal_vae_public_b200 import _lib
    lib = _lib.load()
    rc = lib.mvae_gemm_batch(None, 0, 0, None)
    assert rc == -1 and b"mvae_gemm_batch" in lib.mvae_last_error()
    rc = lib.mvae_colsum_accumulate(None, 0, None, 0, 0, None)
    assert rc == -1
    rc = lib.mvae_gemm_chain(None, None, 0, None, 0, 1, None)
    assert rc == -1 and b"mvae_gemm_chain" in lib.mvae_last_error()
    rc = lib.mvae_allreduce_adam_p2p(None, None, None, None, None, 0, 0, None, 0, 0, 0.0, None, 0.9, 0.999, 1e-8, None, None)
    assert rc == -1 and b"allreduce_adam_p2p" in lib.mvae_last_error()
    with pytest.raises(_lib.MvaeError):
        _lib.check(rc, "colsum")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from multimodal_vae_public_b200 import _lib
    from multimodal_vae_public_b200.trainer import MnistMVAETrainer
    with pytest.raises(_lib.MvaeError):
        MnistMVAETrainer(batch_size=8)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "multimodal_vae_public_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle/"
