"""Time DataCube::Load (our C++ class, 1 and N worker threads) against the reference's loader on the ASCII
export of the 101x161x161 synthetic cube (59 MB).  CPU only:  python tools/load_time.py"""
import json
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import mirror as M  # noqa: E402
import oracle as O  # noqa: E402
from birp_cpp_b200 import synthetic as S  # noqa: E402

zyx, org, sp = S.config1_cube()
path = os.path.join(tempfile.mkdtemp(), "config1.dat")
S.write_dat(path, zyx, org, sp)
mb = os.path.getsize(path) / 1e6
out = {"file_MB": round(mb, 1), "host_threads": os.cpu_count()}


def best(fn, n=5):
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        obj = fn()
        ts.append(time.perf_counter() - t)
        last = obj
    return min(ts), last


for th in ("1", "4", "16"):
    os.environ["BIRP_LOAD_THREADS"] = th
    t, c = best(lambda: M.Cube(path=path))
    out[f"ours_{th}_threads_s"] = round(t, 4)
    got = c.export()
os.environ.pop("BIRP_LOAD_THREADS")
t, c = best(lambda: M.Cube(path=path))
out["ours_default_s"] = round(t, 4)
# side-car: written once (on request), then every Load is one read of the flat grid
M.cache_policy(True)
M.Cube(path=path)
M.cache_policy(False)
t, c = best(lambda: M.Cube(path=path))
out["ours_side_car_s"] = round(t, 4)
out["side_car_used"] = bool(c.state()[1])
out["side_car_MB"] = round(os.path.getsize(path + ".birpcache") / 1e6, 1)
got_cache = c.export()
os.remove(path + ".birpcache")
out["side_car_identical"] = bool(np.array_equal(got_cache.view(np.uint32), got.view(np.uint32)))
if O.have_ref():
    t, r = best(lambda: O.RefCube(path=path), n=3)
    out["reference_s"] = round(t, 4)
    out["identical_to_reference"] = bool(np.array_equal(got.view(np.uint32), r.export().view(np.uint32)))
    out["speedup_default"] = round(out["reference_s"] / out["ours_default_s"], 1)
print(json.dumps(out))
