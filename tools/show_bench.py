"""Print the key fields of a bench.py JSON line (value, e2e, roofline, checks)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("n_gpus", d.get("n_gpus"), "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3),
      "| e2e", round(d["e2e"]["value"], 1), "ms", round(d["e2e"]["ms_per_step"], 3),
      "| roofline", round(d["roofline"]["frac"], 4) if d.get("roofline") else None,
      "| checks", d.get("frame_check"), "| clocks", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
print("parity", d.get("parity"))
print("roofline", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in (d.get("roofline") or {}).items() if k != "timing"})
print("build", d.get("build"))
print("build_soup", d.get("build_soup"))
print("cpu", {k: v for k, v in (d.get("cpu_baseline") or {}).items() if k not in ("sample", "gi")})
