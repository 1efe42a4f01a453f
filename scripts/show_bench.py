"""One-line digest of a bench.py JSON line: python scripts/show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["config"]["workload"], "ms/step", round(d["ms_per_step"], 4), "dev", round(d["value"]), "e2e", round(d["e2e"]["value"]),
      "cycle", round(d["e2e_cycle_api"]["value"]), "cycle_dev_rng", round(d.get("e2e_cycle_device_rng", {}).get("value", 0)),
      "dev ms/step in cycle", round(d["e2e_cycle_api"]["device_ms_per_step"], 4),
      "k_score", d["roofline"]["kernels"]["k_score"]["ms_per_launch"])
