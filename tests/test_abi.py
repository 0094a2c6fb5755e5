"""CPU: the C-ABI shared library loads and exports every symbol include/s2s_b200.h declares.
No compute calls here (no GPU in the builder container)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "s2s_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(s2s_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_listed_in_binding():
    from speech_to_speech_b200 import _lib
    assert sorted(_lib.EXPORTED) == _declared()


def test_library_loads_and_exports_everything():
    from speech_to_speech_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as ge
        ge.build()
    lib = _lib.load()
    for sym in _declared():
        assert hasattr(lib, sym), sym


def test_missing_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from speech_to_speech_b200 import engine
    from speech_to_speech_b200._lib import S2SError
    with pytest.raises(S2SError):
        engine.get_context(0)
