"""Prints the headline numbers of a bench.py JSON line (developer aid).  usage: bench_summary.py <file.json>"""
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][-1]
r = lambda v: round(v, 1) if isinstance(v, float) else v
print("N=%s value %s e2e %s (plain %s, device-inputs %s, link ceiling %s) frac %s cpu %s" % (
    d["n_gpus"], r(d["value"]), r(d["e2e"]["value"]), r(d["e2e"]["plain_calls_value"]), r(d["e2e"]["device_inputs_value"]),
    r(d["e2e"].get("host_link", {}).get("ceiling_problems_per_s")), round(d["roofline"]["frac"], 4), r(d.get("cpu_baseline", {}).get("value"))))
for k, v in d.get("also", {}).items():
    if "e2e" in v:
        cs = v.get("cold_start", {})
        print("  %-16s value %s e2e %s (plain %s) frac %s | cold %s problems/s, %s factorizations, frac %s, status %s" % (
            k, r(v["value"]), r(v["e2e"]["value"]), r(v["e2e"]["plain_calls_value"]), round(v["roofline"]["frac"], 4), r(cs.get("value")),
            r(cs.get("mean_factorizations")), r(cs.get("roofline_frac")), cs.get("status_counts")))
    else:
        print("  %-16s %s" % (k, {a: r(b) for a, b in v.items() if not isinstance(b, (dict, list, str))}))
