import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--emu", action="store_true", default=False,
                     help="development aid: run the C-ABI tests against the host-emulated kernels (tests/emu) instead of the GPU build")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    if config.getoption("--emu"):
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import build_emu
        from tophat_b200 import capi
        capi._lib = capi.load_library(build_emu.build())
        os.environ["THB_TEST_EMU"] = "1"


@pytest.fixture(scope="session")
def built_library():
    from tophat_b200 import build
    return build.build_library()


@pytest.fixture()
def emu_lib(monkeypatch):
    """Development harness (tests/emu): the library's kernels compiled for the host against an emulation shim, so that
    their logic can be checked against the oracle without a GPU.  Test-only: capi's own loader never picks it up."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    from tophat_b200 import capi
    lib = capi.load_library(build_emu.build())
    monkeypatch.setattr(capi, "_lib", lib)
    return lib
