"""Headline numbers of a bench.py JSON line: python scripts/show_bench.py file.json"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
r = d.get("roofline", {})
print("value %.1f e2e %.1f ms/step %.2f n_gpus %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("n_gpus")))
print(" detail", d.get("detail"))
print(" roofline frac %.3f engine %s traffic %s" % (r.get("frac", 0), (r.get("engine_model") or {}).get("frac"), r.get("traffic")))
if "per_iteration_spmv" in r:
    print(" spmv", {k: (round(v["us_per_product"], 1), round(v["frac"], 3)) for k, v in r["per_iteration_spmv"].items() if isinstance(v, dict)})
if "spmv" in d and "variants" in d["spmv"]:
    print(" standalone", {k: round(v["ms_flushed"] * 1e3, 1) for k, v in d["spmv"]["variants"].items()})
for k in ("structured", "batch"):
    if k in d:
        print(" " + k, {a: b for a, b in d[k].items() if a in ("value", "error", "ms_per_step", "ring", "ring_autotune_ms")}, (d[k].get("roofline") or {}).get("frac"), (d[k].get("e2e") or {}).get("value"))
print(" clocks", d.get("clocks"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
