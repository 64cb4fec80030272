import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "osqp-python_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")

# The reference's own test files (tests/ref/suite, and their assembled copy under tests/ref/_work) are run by
# tests/test_refsuite.py in a subprocess with the reference facade on the path; they are not part of this collection.
collect_ignore_glob = ["ref/suite/*", "ref/_work/*"]
