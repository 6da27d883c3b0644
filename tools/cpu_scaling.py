"""Thread scaling of the CPU arm (oracle) on this host: one full-size solve per thread count, with the
per-phase wall times of PSFM_ORACLE_TIMING.  Development aid for bench.py's `--impl reference`."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["PSFM_ORACLE_TIMING"] = "1"
import oracle
from particlesfm_b200 import synthetic as syn, _abi

P = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
prob, _ = syn.make_ba_problem(200, P, 12, seed=5)
o = oracle.ba_global_options(refine_rotation=True, refine_focal_length=True)
o.linear_solver = _abi.SOLVER_AUTO
print("host threads", oracle.num_threads(), flush=True)
for nt in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["8", "16", "32", "64"])]:
    p = prob.copy()
    t0 = time.perf_counter()
    s = oracle.ba_solve(p, o, num_threads=nt)
    dt = time.perf_counter() - t0
    print(f"threads={nt}: {dt:.2f} s, {s.num_iterations} its, {prob.num_observations / dt / 1e6:.3f} M obs/s", flush=True)
