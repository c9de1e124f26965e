"""The asin coefficients of birp_cmem_math.cuh are the ones tools/gen_cmem_coeffs.py derives (mpmath, 60 digits):
a table in a header is only as trustworthy as the recipe that reproduces it."""
import re
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_asin_table_is_reproducible():
    pytest.importorskip("mpmath")
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "gen_cmem_coeffs.py"), "10"], capture_output=True,
                         text=True, check=True).stdout
    generated = [float.fromhex(h) for h in re.findall(r"(-?0x1\.[0-9a-f]+p[+-]?\d+)", out)]
    assert len(generated) == 11  # z^0 .. z^10
    err = float(re.search(r"max abs error ([0-9.e+-]+)", out).group(1))
    assert err < 1e-15
    header = (ROOT / "birp_cpp_b200" / "csrc" / "birp_cmem_math.cuh").read_text()
    block = header[header.index("kAsinC[11]"):]
    block = block[:block.index("};")]
    table = [float.fromhex(h) for h in re.findall(r"(-?0x1\.[0-9a-f]+p[+-]?\d+)", block)]
    assert table == generated[::-1], "header table (highest power first) != generator output"
