"""CPU: the C-ABI library loads and exports every symbol include/add_b200.h declares; argument
validation returns status codes (no compute is launched without a GPU)."""
import ctypes
import re

import util
import add_b200
from add_b200 import _lib

HEADER = (util.ROOT / "include" / "add_b200.h").read_text()


def declared_symbols():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(add_[a-z0-9_]+)\s*\(", body)))


def test_every_declared_symbol_is_exported():
    names = declared_symbols()
    assert len(names) >= 20
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/add_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), set(names) ^ set(_lib.EXPORTED_SYMBOLS)


def test_version_and_status_strings():
    assert _lib.lib.add_version() >= 100
    assert _lib.lib.add_status_string(0) == b"ok"
    for code in (-1, -2, -3, -4):
        assert len(_lib.lib.add_status_string(code)) > 3


def test_bad_arguments_return_status_not_crash():
    t = _lib.AddTensor(None, 1, 4, 4, 8, 8, 0)
    rc = _lib.lib.add_bilinear_fwd(ctypes.byref(t), ctypes.byref(t), 0, None)
    assert rc == -1
    assert _lib.lib.add_confusion_workspace_bytes(-5, 19) == -1
    assert _lib.lib.add_head_workspace_bytes(0, 4, 4, 19) == -1
    assert _lib.lib.add_confusion_workspace_bytes(1000, 19) > 0


def test_no_cpu_fallback():
    import pytest
    import torch
    m, x = util.make_op_case("sep_conv_3x3_c40")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x)
    net = util.make_net(util.NET_CASES["searched-dense-C2"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(torch.zeros(1, 3, 33, 65))
    ev = add_b200.Evaluator(19)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ev.add_batch(torch.zeros(1, 4, 4, dtype=torch.int64), torch.zeros(1, 4, 4, dtype=torch.int64))
