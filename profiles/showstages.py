import json, sys
for n in sys.argv[1:]:
    j = json.loads(open(n).read().strip().splitlines()[-1])
    print(n, "%.2f ms/step, e2e %.2f" % (j["ms_per_step"], j["e2e"]["ms_per_step"]))
    print("  ", j["roofline"]["stage_ms_per_step"])
    print("  ", {k: v for k, v in j["work"].items() if k in ("pairs", "mid_reads", "slow_reads", "em_classes", "em_class_pairs")})
    for name, c in (j.get("configs") or {}).items():
        if "points" in c:
            for s, p in c["points"].items():
                if "ms_per_step" in p:
                    print("  ", name, s, "%.2f ms/step" % p["ms_per_step"], p["stage_ms_per_step"])
        elif "ms_per_step" in c:
            print("  ", name, "%.2f ms/step" % c["ms_per_step"], c["stage_ms_per_step"], {k: v for k, v in c["work"].items() if k in ("mid_reads", "slow_reads")})
