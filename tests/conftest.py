import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))     # the oracle is test infrastructure: only tests import it

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_sessionstart(session):
    """The CPU tier checks the ABI surface of the built library and the C++ command line: build them if this
    checkout has not been built yet (nvcc cross-compiles without a GPU; on the GPU box the built files travel)."""
    lib = os.path.join(ROOT, "fastani_b200", "lib", "libfastani_b200.so")
    exe = os.path.join(ROOT, "fastani_b200", "bin", "fastANI")
    if not (os.path.exists(lib) and os.path.exists(exe)):
        from fastani_b200 import build
        build.build()
        if not os.path.exists(exe):
            build.build_cli()
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(orc):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], stdout=subprocess.DEVNULL)
