"""Application-level timing (not a test): the reference's icoFoam on an n^3 lid-driven cavity, a few time steps,
(a) with the reference's own solvers on one host core, (b) with libgpuLduSolvers.so picked up from system/controlDict
(LDU_GPU_OVERRIDE=1, unmodified fvSolution).  usage: perf_icofoam.py [n] [steps]"""
import re
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
import foam_case as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dt = 0.05 / n            # Courant number ~0.5 at lid speed 1 on cells of 0.1/n
LINE = re.compile(r"(\S+):\s+Solving for (\w+), Initial residual = (\S+), Final residual = (\S+), No Iterations (\d+)")


def summary(log):
    its = {}
    for x in F.solver_lines(log):
        m = LINE.match(x)
        its.setdefault(m.group(2), []).append(int(m.group(5)))
    clock = [float(x.split("ClockTime =")[1].split()[0]) for x in log.splitlines() if "ClockTime" in x]
    return its, (clock[-1] if clock else None)


with tempfile.TemporaryDirectory() as td:
    kw = dict(nx=n, ny=n, nz=n, end_time=steps * dt, delta_t=dt)
    t0 = time.perf_counter()
    cpu = F.write_cavity(Path(td) / "cpu", **kw)
    F.run("blockMesh", cpu)
    gpu = F.write_cavity(Path(td) / "gpu", libs=[str(F.PLUGIN)], **kw)
    F.run("blockMesh", gpu)
    print(f"cavity {n}^3 = {n**3} cells, {steps} time steps, blockMesh x2 {time.perf_counter() - t0:.1f} s", flush=True)
    t0 = time.perf_counter()
    log_gpu = F.run("icoFoam", gpu, env=dict(LDU_GPU_OVERRIDE="1"))
    t_gpu = time.perf_counter() - t0
    its_gpu, _ = summary(log_gpu)
    print(f"GPU plug-in : icoFoam wall {t_gpu:.1f} s; iterations p {its_gpu.get('p')} Ux {its_gpu.get('Ux')}", flush=True)
    t0 = time.perf_counter()
    log_cpu = F.run("icoFoam", cpu, timeout=7200)
    t_cpu = time.perf_counter() - t0
    its_cpu, _ = summary(log_cpu)
    print(f"reference   : icoFoam wall {t_cpu:.1f} s; iterations p {its_cpu.get('p')} Ux {its_cpu.get('Ux')}")
    print(f"same iteration counts: {its_cpu == its_gpu};  application speed-up {t_cpu / t_gpu:.1f}x "
          f"(whole icoFoam incl. start-up, assembly on the CPU in both runs)")
