"""One-line-per-field summary of a bench.py JSON line (development aid)."""
import json, sys
j = json.load(open(sys.argv[1]))
print("value %.4e  ms/step %.4f  roofline %.3f (kernel %.4f ms)  host enqueue %.4f ms" % (
    j["value"], j["ms_per_step"], j["roofline"]["frac"], j["roofline"]["kernel_ms"], j.get("host_enqueue_ms_per_step") or 0))
for k in ("e2e", "e2e_resident", "e2e_device_loop"):
    v = j.get(k)
    if v:
        print("  %-16s %.4e  %s" % (k, v["value"], {x: v[x] for x in v if x.startswith("ms_") or x in ("attempts", "accepted")}))
for k, v in (j.get("also") or {}).items():
    print("  also.%s  %.4e  ms/step %.4f  roofline %.3f" % (k, v["value"], v["ms_per_step"], v["roofline"]["frac"]))
if j.get("cpu_baseline"):
    print("  cpu_baseline %.4e (%d cores)" % (j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"]))
print("  clocks", j.get("clocks"))
