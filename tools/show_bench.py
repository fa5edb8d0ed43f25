"""One-screen digest of a bench.py JSON line:  python tools/show_bench.py <file.json>"""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)      # quiet under `| head`

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"]), d["unit"], "| ms/step", round(d["ms_per_step"], 1), "| e2e", round(d["e2e"]["value"]), "| gpus", d["n_gpus"],
      "| launches", d.get("gpu_launches"))
print("pcg", d.get("pcg"))
r = d.get("roofline") or {}
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ("achieved", "frac", "avg_launch_ms", "share_of_step", "traffic", "launches_timed")})
print("kernels", {k: {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()} for k, v in (d.get("kernels") or {}).items()})
print("phases", {k: round(v, 2) for k, v in (d.get("phases_ms_per_step") or {}).items()}, "| cpu", (d.get("cpu_baseline") or {}).get("value"), "| clocks", d.get("clocks"))
