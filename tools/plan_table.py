#!/usr/bin/env python
"""Per-layer table of a net: the host-only launch plan (b200_fwd_plan) beside the per-call eager profile a bench line carries.
usage: plan_table.py NET BATCH PREC bench_line.json > table.md      (the bench line must come from `bench.py --net NET --batch BATCH --prec PREC`)
Eager per-call times carry a ~8-10 us launch + event floor per call; the graph replay total is the line's ms_per_step."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import boda_b200 as bb
from boda_b200 import nets


def main():
    net, batch, prec, fn = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    d = json.load(open(fn))
    txt, i, o = nets.NETS[net](batch)
    plan = bb.fwd_plan(txt, "(prec=%s)" % prec)
    prof = {r["func"]: r for r in d["per_call"]}
    print("# %s B=%d %s: launch plan and eager per-call profile\n" % (net, batch, prec))
    print("Graph replay: %.4f ms/step (%.0f images/s). Sum of eager calls: %.3f ms. Plans from `b200_fwd_plan` (host-only), times from `%s`.\n"
          % (d["ms_per_step"], d["value"], sum(r["call_ms"] for r in d["per_call"]), os.path.basename(fn)))
    print("| call | kernel | bn | k-blocks | splits | grid | GFLOP | call us | contraction us | TF/s (contraction) | extra args |\n|---|---|---|---|---|---|---|---|---|---|---|")
    for f, a in plan["calls"]:
        r = prof.get(f)
        p = {k[5:]: v for k, v in a.items() if k.startswith("plan:")}
        extra = ",".join(k for k in ("res", "out_concat", "out_pack") if k in a)
        if r is None:
            continue
        gf = float(r["gflop"])
        kus = 1e3 * float(r["kernel_ms"])
        print("| %s | %s | %s | %s | %s | %s | %s | %.1f | %.1f | %s | %s |" % (f.split("__", 1)[0] + " " + f.split("__")[1], p.get("kernel", ""), p.get("bn", ""), p.get("kblks", ""),
              p.get("splits", ""), p.get("grid", ""), ("%.2f" % gf) if gf else "", 1e3 * float(r["call_ms"]), kus, ("%.0f" % (gf / kus * 1e3)) if gf else "", extra))


if __name__ == "__main__":
    main()
