"""CPU checks of the C-ABI boundary: libspsk.so loads without a GPU, exports every symbol include/spsk.h
declares, and rejects bad arguments with a status code + message instead of exiting (no compute calls)."""
import ctypes as C
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_exports_every_declared_symbol():
    from spsnet_b200 import _lib

    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(_lib.lib, name), f"libspsk.so does not export {name}"
    assert set(_lib.SIGNATURES) == set(declared), "ctypes signature table out of sync with include/spsk.h"


def test_header_is_pure_c_abi():
    txt = (ROOT / "include" / "spsk.h").read_text()
    assert 'extern "C"' in txt
    assert "torch" not in re.sub(r"/\*.*?\*/", "", txt, flags=re.S), "no torch types may appear in the C-ABI"
    for ref in ("src/sampling.cpp:34-43", "src/ball_query.cpp:32-42", "src/group_points.cpp:30-40", "src/interpolate.cpp:21-30"):
        assert ref in txt, f"header must cite the reference interface it replaces ({ref})"


def test_identity_and_errors():
    from spsnet_b200 import _lib

    lib = _lib.lib
    assert lib.spsk_abi_version() == 5
    assert lib.spsk_built_for_sm() == 100
    assert lib.spsk_ball_query(1, 8, 4, 1.0, 0, None, None, None, None) == -1
    assert b"null" in lib.spsk_last_error() or b"nsample" in lib.spsk_last_error()
    assert lib.spsk_farthest_point_sampling(-1, 4, 2, None, None, None, None) == -1
    assert lib.spsk_score_topk(1, 100000, 3, 10, None, None, None, None, None) in (-1, -2)
    assert lib.spsk_score_topk(1, 10, 3, 20, None, None, None, None, None) == -1  # npoint > n
    try:
        _lib.check(-2, "x")
    except _lib.SpskError as e:
        assert "status -2" in str(e)
    else:
        raise AssertionError("check() must raise")
    # empty problems are no-ops that succeed without touching the device
    assert lib.spsk_gather_points(0, 4, 10, 5, None, None, None, None) == 0
    assert lib.spsk_three_nn(0, 0, 0, None, None, None, None, None) == 0


def test_no_cpu_fallback_in_product():
    """The product package must not import the oracle (parity claims are void otherwise)."""
    for py in (ROOT / "spsnet_b200").glob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f"{py.name} references the oracle"
