import json,sys
for n in sys.argv[1:]:
    try:
        d=json.load(open(n))
    except Exception as e:
        print(n, "unreadable", e); continue
    print(n, "%.4g reads/s %.2f ms | e2e %.4g %.2f ms | %s frac %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel"], d["roofline"]["frac"]))
    print("  ", d["roofline"]["stage_ms_per_step"]); print("  ", {k:v for k,v in d["work"].items() if k in ("mid_reads","slow_reads","overflow_reads","queries","hits")})
