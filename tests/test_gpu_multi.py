"""Multi-region (one process per region) runs of the CUDA path: processor-patch
halos and scalar all-reduces through the peer-mapped exchange window, checked
against the oracle's in-process world of the same regions."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(world, n=12, timeout=600):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LDU_PORT=str(port), LDU_N=str(n))
        procs.append(subprocess.Popen([sys.executable, str(ROOT / "tests" / "multi_rank_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            outs.append(p.communicate(timeout=timeout)[0])
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
    res = []
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-4000:]}"
        line = [x for x in o.splitlines() if x.startswith("RESULT ")][-1]
        res.append(json.loads(line[7:]))
    return res


@pytest.mark.parametrize("world", [2, 4])
def test_multi_region_parity(world):
    for rank, r in enumerate(_run(world)):
        assert r["amul"] and r["residual"] and r["sumA"] and r["gs"], (rank, r)
        for name in ("symGaussSeidel", "nonBlockingGaussSeidel", "DIC", "FDIC", "DICGaussSeidel",
                     "multiColourGaussSeidel"):
            assert r["smooth_" + name], (rank, name, r)
        assert r["gamg_nbgs"] and r["gamg_mcgs"], (rank, r)
        for i in range(3):
            it, ito = r[f"solve{i}_exact_iters"]
            assert it == ito, (rank, i, r)
            res, reso = r[f"solve{i}_exact_res"]
            assert res == reso, (rank, i, r)                 # reference-order sums: bit-identical
            assert r[f"solve{i}_exact_psi"] == 0.0, (rank, i, r)
            it, ito = r[f"solve{i}_fast_iters"]
            assert it == ito, (rank, i, r)
            res, reso = r[f"solve{i}_fast_res"]
            assert abs(res - reso) <= 1e-4 * reso + 1e-13
