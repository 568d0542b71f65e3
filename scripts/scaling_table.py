"""Runs bench.py for every BASELINE config at N GPUs (N = number of visible GPUs) at a reduced spp
and appends the JSON lines to gpurun_out/scaling_N.jsonl.   python scripts/scaling_table.py <N> [spp]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]); spp = sys.argv[2] if len(sys.argv) > 2 else "256"
out = os.path.join(ROOT, "gpurun_out", f"scaling_{n}.jsonl")
os.makedirs(os.path.dirname(out), exist_ok=True)
open(out, "w").close()
only = os.environ.get("ONLY", "C1,C2,C3,C4,C4b,C5").split(",")
for i, wl in enumerate(only):
    cmd = ["bench.py", "--gpus", str(n), "--steps", "2", "--warmup", "3", "--workload", wl, "--spp", spp, "--no-cpu-baseline"]
    if n > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
               "--master-port", str(29600 + i)] + cmd
    else:
        cmd = [sys.executable] + cmd
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    if not lines:
        print(wl, "FAILED", res.stderr[-800:])
        continue
    d = json.loads(lines[-1])
    with open(out, "a") as f:
        f.write(lines[-1] + "\n")
    print(wl, n, "GPUs:", round(d["value"]), "Mpaths/s,", round(d["mrays_per_s"]), "Mrays/s, ms/frame", round(d["ms_per_step"], 2),
          "e2e", round(d["e2e"]["value"]), "Mpaths/s (", round(d["e2e"]["ms_per_step"], 1), "ms ) roofline frac", round(d["roofline"]["frac"], 3), flush=True)
