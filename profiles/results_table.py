"""Prints the markdown tables of DESIGN.md section 8 from the committed bench lines under profiles/.
python profiles/results_table.py"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with open(name if os.path.isabs(name) or os.path.isfile(name) else os.path.join(HERE, name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def fmt(v):
    return "%.3g" % v if v < 1000 else "{:,.0f}".format(v).replace(",", " ")


def main():
    import sys
    d = load(sys.argv[1] if len(sys.argv) > 1 else "r2_bench_n1.json")
    ref = load("r2_bench_reference_n1.json")
    rows = [("configs[4] BPRMF gd, 6.25 M x 12.5 M x 128 (headline)", d, ref["value"], ref["cpu_baseline"]["cores"])]
    for k in ("bprmf-ml100k", "neumf-ml100k", "lightgcn-gowalla"):
        o = d["others"][k]
        rows.append((k, o, o["cpu_baseline"]["value"], o["cpu_baseline"]["cores"]))
    print("| workload (N=1) | value | e2e | ms/step | CPU arm (threads) | value / CPU | e2e / CPU | roofline kernel | frac |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, o, cpu, cores in rows:
        r = o["roofline"]
        print("| %s | %s %s | %s | %.4g | %s (%d) | %.0fx | %.0fx | `%s` %s %.0f / %.0f %s | %.3f |" % (
            name, fmt(o["value"]), o["unit"], fmt(o["e2e"]["value"]), o["ms_per_step"], fmt(cpu), cores, o["value"] / cpu,
            o["e2e"]["value"] / cpu, r["kernel"], r["bound"], r["achieved"], r["peak"], r["unit"], r["frac"]))
    ev = d["others"]["eval-synth"]
    r = ev["roofline"]
    print("| eval-synth (configs[3]) | %s %s | %s | %.4g | %s (%d) | %.0fx | %.0fx | `%s` %s %.0f / %.0f %s | %.3f |" % (
        fmt(ev["value"]), ev["unit"], fmt(ev["e2e"]["value"]), ev["ms_per_step"], fmt(ev["cpu_baseline"]["value"]),
        ev["cpu_baseline"]["cores"], ev["value"] / ev["cpu_baseline"]["value"], ev["e2e"]["value"] / ev["cpu_baseline"]["value"],
        r["kernel"], r["bound"], r["achieved"], r["peak"], r["unit"], r["frac"]))
    print()
    print("| evaluator (N=1) | users/s | CPU users/s (threads) | ratio | NDCG@10 |")
    print("|---|---|---|---|---|")
    for k in ("bprmf-ml100k", "lightgcn-gowalla"):
        e = d["others"][k]["eval"]
        print("| %s: %d users x %d items | %s | %s (%d) | %.0fx | %.6f |" % (k, e["users"], e["items"], fmt(e["value"]),
                                                                          fmt(e["cpu"]["value"]), e["cpu"]["threads"],
                                                                          e["value"] / e["cpu"]["value"], e["ndcg_at_10"]))
    la = d["lazy_adam"]
    print()
    print("lazy-Adam run: %s triplets/s, %.3f of the HBM peak (%s)" % (fmt(la["value"]), la["roofline"]["frac"], la["roofline"]["kernel"]))
    print()
    print("| N | G triplets/s | per GPU | vs N=1 per GPU | kernel us | head sync us | NVLink GB/s per GPU and direction | of 770 | e2e G/s | evaluator users/s (configs[3], users sharded) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    base = d["value"]
    for n, name in ((1, "r2_bench_n1.json"), (2, "r2_bench_n2.json"), (4, "r2_bench_n4.json"), (8, "r2_bench_n8.json")):
        if not os.path.isfile(os.path.join(HERE, name)):
            continue
        x = load(name)
        r = x["roofline"]
        nv = r.get("nvlink", {})
        ev = x.get("others", {}).get("eval-synth")
        print("| %d | %.3f | %.3f | %.2f | %.0f | %.0f | %s | %s | %.3f | %s |" % (
            n, x["value"] / 1e9, x["value"] / 1e9 / n, x["value"] / n / base, r["launch_us"], r["replicated_head"]["sync_us_mean"],
            "%.0f" % nv["GBps_per_gpu_per_direction"] if nv else "-", "%.2f" % nv["of_measured_peer_copy_770_GBps"] if nv else "-",
            x["e2e"]["value"] / 1e9, fmt(ev["value"]) if ev else "(line predates the evaluator entry)"))


if __name__ == "__main__":
    main()
