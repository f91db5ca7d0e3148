import json, sys
lines = [l for l in sys.stdin.read().splitlines() if l.startswith('{')]
name = sys.argv[1] if len(sys.argv) > 1 else ""
if not lines:
    print(name, "NO JSON LINE")
    sys.exit(0)
d = json.loads(lines[-1])
print(name, "ms/step", round(d["ms_per_step"], 4), "Gcells/s", round(d["value"] / 1e9, 3),
      "frac", round(d["roofline"]["step_frac"], 3), d["roofline"]["kernels_ms"], "GiB", round(d["device_gib"], 1),
      "graph_ms", d["config"].get("graph_ms_per_step"), "e2e_ms", round(d["e2e"]["ms_per_step"], 3))
