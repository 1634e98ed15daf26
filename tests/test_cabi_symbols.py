"""CPU: the C-ABI library loads and exports every symbol include/cv2eu_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, "include", "cv2eu_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cv2_\w+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from cosyvoice2_eu_b200 import lib
    assert os.path.exists(lib.LIB_PATH), "run __graft_entry__.build() first"
    dll = ctypes.CDLL(lib.LIB_PATH)
    decl = _declared()
    assert len(decl) >= 18
    for name in decl:
        assert hasattr(dll, name), name
    assert sorted(lib.SYMBOLS) == decl


def test_no_cpu_fallback_without_gpu():
    """The product path fails loudly when no CUDA device is present."""
    import pytest
    import torch
    from cosyvoice2_eu_b200 import lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cosyvoice2_eu_b200 import B200Flow
    with pytest.raises(lib.Cv2Error):
        B200Flow("cuda:0")
    L = lib.load()
    h = ctypes.c_void_p()
    assert L.cv2_engine_create(ctypes.byref(h), 0) != 0
    assert b"no CUDA device" in L.cv2_last_error() or b"fallback" in L.cv2_last_error()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cosyvoice2_eu_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_header_is_plain_c(tmp_path):
    """include/cv2eu_b200.h must be consumable by a C compiler as is (the drop-in boundary is a C ABI: plain pointers and sizes)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no gcc")
    src = tmp_path / "use_header.c"
    src.write_text('#include "cv2eu_b200.h"\n'
                   "int probe(void) {\n"
                   "  cv2_engine* e = 0;\n"
                   "  size_t n = cv2_prompt_mel_workspace_bytes(1, 24000);\n"
                   "  return cv2_version() + (int)n + (e != 0) + cv2_resample_16k_24k_len(16000) + cv2_prompt_mel_frames(24000);\n"
                   "}\n")
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
