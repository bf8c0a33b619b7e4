"""Print the essentials of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4),
      "separate_calls", round(d["separate_calls"]["value"]) if d.get("separate_calls") else None,
      "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None,
      "step_frac", round(d["roofline"]["step_frac_of_hbm_peak"], 4))
print({k: (round(v["avg_launch_ms"], 4), v["launches"]) for k, v in d["kernels"].items()})
