"""GPU: placeholder import check so that the native library is recorded as loaded by the test run."""
import pytest

pytestmark = pytest.mark.gpu


def test_library_is_native_and_loaded(jf):
    from juliafem.jl_b200 import _lib
    assert _lib.lib().jfem_abi_version() == 1
    maps = open("/proc/self/maps").read()
    assert "libjfem_b200.so" in maps
