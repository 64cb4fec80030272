"""Host logic of the tile format (csrc/csl_host.h), no GPU: the builder's blobs decode back to the CSR they were cut
from (two-part n-side tiles and one-part m-side tiles, ragged and empty rows, row counts around the 32-row tile size),
the value maps hit every entry, tiles start on 128-byte lines, the CTA / rank partitions are monotone covers, and a
row longer than 255 entries is refused (the engine then keeps its CSR kernel)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_builder_round_trip_and_partitions(tmp_path):
    exe = str(tmp_path / "csl_host_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "osqp-python_b200", "csrc"),
                           os.path.join(ROOT, "tests", "host", "csl_host_check.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "csl_host OK" in out.stdout, out.stdout + out.stderr
