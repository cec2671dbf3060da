import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
"""Print the per-level table of a bench JSON."""
import json, sys
d = json.load(open(sys.argv[1]))
print({k: v for k, v in d.items() if k not in ("levels", "config", "cpu_baseline", "clocks")})
print("clocks", d.get("clocks"))
for l in d["levels"]:
    print(f"{l['level']:10s} {l['ms']*1e3:8.1f} us  roof {l['roof_ms']*1e3:6.1f} us  {l['bound']:6s} {l['gbs']:7.1f} GB/s {l['tflops']:7.1f} TF  frac {l['frac']:.3f}")
