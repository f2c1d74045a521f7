import json,sys
for f in sys.argv[1:]:
    d=json.load(open(f)); r=d["roofline"]; v=r.get("vjp_kernel", r)
    print(f, "%.1f G c-s/s  step %.3f ms | dominant frac %.3f | vjp %.3f ms frac %.3f | rhs %.3f ms frac %.3f | e2e %.2f G" % (d["value"]/1e9, d["ms_per_step"], r["frac"], v["ms_per_launch"], v["frac"], r["rhs_kernel"]["ms_per_launch"], r["rhs_kernel"]["frac"], (d["e2e"]["value"] or 0)/1e9))
