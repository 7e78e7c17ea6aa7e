import json, sys
d = json.load(open(sys.argv[1]))
print("ms/step", round(d["ms_per_step"], 4), "e2e", round(d.get("e2e", {}).get("ms_per_step", 0), 4),
      {k: round(v, 4) for k, v in d["roofline"]["kernels_ms_per_step"].items()}, "value", f"{d['value']:.4g}")
