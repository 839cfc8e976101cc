import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs the reference build under oracle/_ref")


@pytest.fixture(scope="session")
def cpp_tool(tmp_path_factory):
    """Builds tests/cpp/<name>.cpp once per session (g++ -O2 -pthread -lz) and returns the executable's path."""
    import subprocess
    built = {}

    def build(name):
        if name not in built:
            exe = str(tmp_path_factory.mktemp("cpp_" + name) / name)
            subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
                            "-lz", "-o", exe], check=True)
            built[name] = exe
        return built[name]

    return build
