"""pytest configuration: registers the `gpu` marker and puts the repo's helper
directories on sys.path (the package directory name contains a hyphen, so it is
not importable as a dotted module)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "cmsis-dsp_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")
