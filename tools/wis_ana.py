#!/usr/bin/env python
"""wis-ana for this repo: per-op times measured with be=b200 beside the reference's published database.

The analysis itself is the C++ restatement of the reference's wis-ana mode (b200_wis_ana, boda_b200/csrc/wisdom.cu; src/op-tuner.cc:204-396);
this script joins two passes of it by op text -- the reference's database filtered to one platform (AOM = its best single tune, POM = per-op
best tune, REF = its cuDNN tune) and wisdom files written by tools/ops_prof.py --wisdom-out for be=b200 -- and prints a markdown table.

  python tools/wis_ana.py --ref-wisdom tests/golden/wisdom_merged_c3_subset.wis --b200-wisdom fp16=profiles/wisdom_r01_c3_fp16.wis \\
      --b200-prof fp32=profiles/ops_prof_c3_fp32.json:ops/c3-conv-ops-small.txt --b200-prof bf16=profiles/ops_prof_c3_bf16.json:ops/c3-conv-ops-small.txt \\
      --s-img 20 --md-out profiles/wis_ana_r01_c3.md --csv-out profiles/wis_ana_r01_c3_titanx.csv
(--ref-wisdom may also be the reference's full test/wisdom-merged.wis.)"""
import argparse, json, math, os, sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import boda_b200 as bb

CUDNN_TUNE = "(use_be=nvrtc,use_culibs=1,MNt=8 8,MNb=8 16,tconv_max_ksz=11 11)"


def short_op(op_text):
    import re
    g = lambda pat: re.search(pat, op_text)
    f = g(r"filts=\(dims=\(out_chan=(\d+),in_chan=(\d+),y=(\d+),x=(\d+)\)\)")
    i = g(r"in=\(dims=\(img=(\d+),chan=(\d+),y=(\d+),x=(\d+)\)\)")
    s = g(r"stride=\(tn=none,dims=\(y=(\d+),x=(\d+)\)\)")
    p = g(r"in_pad=\(tn=none,dims=\(y=(\d+),x=(\d+)\)\)")
    if not (f and i and s and p):
        return op_text[:60]
    return "%sx%s/%s/%s %s->%s @%sx%s B=%s" % (f.group(3), f.group(4), s.group(1), p.group(1), f.group(2), f.group(1), i.group(3), i.group(4), i.group(1))


def prof_json_to_wisdom(json_fn, ops_fn, tag):
    """ops_prof JSON rows are in the order of the ops file: rebuild wisdom records (kernel time) for them."""
    rows = json.load(open(json_fn))
    ops = [l.strip() for l in open(ops_fn) if l.strip()]
    if len(rows) != len(ops):
        raise SystemExit("%s has %d rows, %s has %d ops" % (json_fn, len(rows), ops_fn, len(ops)))
    return "".join(bb.wisdom_record(op, [], "(use_be=b200,prec=%s)" % tag, "b200:NVIDIA B200", float(r["kernel_ms"]) * 1e-3) for op, r in zip(ops, rows))


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--ref-wisdom", required=True)
    ap.add_argument("--ref-plat", default="nvrtc:GeForce GTX TITAN X", help="regex over the platform tag of the database runs")
    ap.add_argument("--ref-tune", default=CUDNN_TUNE, help="the tune reported as REF (the reference's cuDNN tune)")
    ap.add_argument("--b200-wisdom", action="append", default=[], metavar="TAG=FILE")
    ap.add_argument("--b200-prof", action="append", default=[], metavar="TAG=JSON:OPSFN")
    ap.add_argument("--s-img", type=int, default=0)
    ap.add_argument("--min-flops", type=float, default=0.0)
    ap.add_argument("--csv-out", default="", help="the reference-format csv of the database pass (input of its wis-plot.py)")
    ap.add_argument("--md-out", default="")
    args = ap.parse_args()

    ref_text = open(args.ref_wisdom).read()
    ref = bb.wis_ana(ref_text, s_img=args.s_img, s_plat=args.ref_plat, ref_tune=args.ref_tune, min_flops=args.min_flops)
    if args.csv_out:
        open(args.csv_out, "w").write(bb.wis_ana(ref_text, s_img=args.s_img, s_plat=args.ref_plat, ref_tune=args.ref_tune, min_flops=args.min_flops, csv=True))
    mine = []
    for spec in args.b200_wisdom:
        tag, fn = spec.split("=", 1)
        mine.append((tag, bb.wis_ana(open(fn).read(), s_img=args.s_img, s_plat="b200:", min_flops=args.min_flops)))
    for spec in args.b200_prof:
        tag, rest = spec.split("=", 1)
        jfn, ofn = rest.split(":", 1)
        mine.append((tag, bb.wis_ana(prof_json_to_wisdom(jfn, ofn, tag), s_img=args.s_img, s_plat="b200:", min_flops=args.min_flops)))
    by_op = {tag: {r["op"]: r["pom"] for r in res["rows"]} for tag, res in mine}

    out = []
    out.append("# wis-ana: be=b200 per-op times beside the reference's database (%s, B=%s)\n" % (args.ref_plat, args.s_img or "all"))
    out.append("Database: `%s` (%d runs used; best single Boda tune = `%s`; REF tune = `%s`). Times in ms; TF/s = algorithmic FLOPs (2*M*N*K) / time. "
               "b200 columns: contraction-kernel time per precision mode (operand packs not included, as the database's times exclude the reference's xpose kernels).\n"
               % (args.ref_wisdom, ref["tot_runs"], ref["aom_tune"], args.ref_tune))
    tags = [t for t, _ in mine]
    out.append("| op | GFLOP | Titan X best single tune | Titan X per-op best | Titan X cuDNN (REF) | " + " | ".join("B200 %s" % t for t in tags) + " | " +
               " | ".join("x cuDNN-TitanX (%s)" % t for t in tags) + " |")
    out.append("|---|---|---|---|---|" + "---|" * (2 * len(tags)))
    fmt = lambda secs, fl: "-" if (secs is None or math.isnan(secs)) else "%.3f (%.2f)" % (secs * 1e3, fl / secs / 1e12)
    for r in sorted(ref["rows"], key=lambda r: r["flops"]):
        b = [by_op[t].get(r["op"]) for t in tags]
        if not any(x is not None for x in b):
            continue
        sp = ["-" if (x is None or math.isnan(r["ref"])) else "%.0fx" % (r["ref"] / x) for x in b]
        out.append("| %s | %.2f | %s | %s | %s | %s | %s |" % (short_op(r["op"]), r["flops"] / 1e9, fmt(r["aom"], r["flops"]), fmt(r["pom"], r["flops"]), fmt(r["ref"], r["flops"]),
                                                          " | ".join(fmt(x, r["flops"]) for x in b), " | ".join(sp)))
    text = "\n".join(out) + "\n"
    if args.md_out:
        open(args.md_out, "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
