import json, sys
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print(sys.argv[1] if len(sys.argv) > 1 else "", "ms/step", round(d["ms_per_step"], 3), "Gcells/s", round(d["value"] / 1e9, 3),
      "frac", round(d["roofline"]["step_frac"], 3), d["roofline"]["kernels_ms"], "GiB", round(d["device_gib"], 1))
