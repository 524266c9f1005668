"""CPU: the C-ABI library loads and exports every symbol include/rgie.h declares (no compute calls without a GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "rgie.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rgie_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = _declared()
    for must in ("rgie_filter_fwd", "rgie_filter_bwd", "rgie_resize_fwd", "rgie_resize_bwd", "rgie_regressor_create",
                 "rgie_regressor_forward", "rgie_regressor_backward", "rgie_va_head", "rgie_adam_step",
                 "rgie_guidance_update", "rgie_midu_forward", "rgie_midu_backward", "rgie_last_error"):
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from regressor_guided_image_editing_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), f"librgie.so does not export {name}"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert lib.rgie_version() == 1
    assert lib.rgie_filter_param_count(_lib.F_COLOR) == 24


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from regressor_guided_image_editing_b200 import _lib, ops
    from regressor_guided_image_editing_b200.baselines.image_transformations import image_transformations as IT
    with pytest.raises(_lib.RgieError):
        IT.apply_exposure(torch.rand(1, 3, 8, 8), torch.tensor(0.1))
    with pytest.raises(_lib.RgieError):
        ops.Regressor({"fc.weight": torch.zeros(4, 2048)}, max_crops=10)


def test_no_oracle_import_in_product_package():
    pkg = os.path.join(ROOT, "regressor_guided_image_editing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} imports the oracle"
