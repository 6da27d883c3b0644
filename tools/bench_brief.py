"""python tools/bench_brief.py [bench.py args]: runs bench.py (no CPU baseline, no HP1, no e2e) and prints the
step time and the per-kernel launch times only — for A/B runs under environment switches."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-traj", "--no-e2e"] + sys.argv[1:],
                     capture_output=True, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print(out.stdout[-2000:], out.stderr[-2000:])
    sys.exit(1)
for l in out.stderr.splitlines():
    if l.startswith("[psfm"):
        print(l)
        break
d = json.loads(line[-1])
env = {k: v for k, v in os.environ.items() if k.startswith("PSFM_")}
print(env, f"ms_per_step {d['ms_per_step']:.3f} its {d['lm_iterations_per_step']} final_cost {d['final_cost']:.6f} units {d.get('pair_units')}")
for k in d.get("roofline_all_kernels", []):
    print(f"   {k['kernel'][:60]:60s} {k['avg_launch_ms']:.4f} ms  share {k['share_of_step']:.3f}")
