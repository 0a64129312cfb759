#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` dump of one hot-path forward (see capture_forward.sh):
    python profiles/summarize_forward.py gpurun_out/r1f_raw.csv profiles/r1f
writes <prefix>_ncu_forward.json (one record per launch), <prefix>_ncu_traffic.json (per kernel: DRAM bytes, time,
tensor-pipe % -- what bench.py reads for `roofline.traffic`) and prints the per-kernel markdown table."""
import collections
import csv
import json
import sys


def main():
    src, prefix = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def g(r, k):
        try:
            return float(r[idx[k]].replace(",", ""))
        except Exception:
            return 0.0

    mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rmul = mul.get(units[idx["dram__bytes_read.sum"]], 1.0)
    wmul = mul.get(units[idx["dram__bytes_write.sum"]], 1.0)
    tmul = {"us": 1.0, "ns": 1e-3, "ms": 1e3}.get(units[idx["gpu__time_duration.sum"]], 1.0)
    out, traffic = [], collections.defaultdict(list)
    body = rows[2:]
    if len(sys.argv) > 3 and sys.argv[3] == "--second-half":      # two forwards captured: keep the second (warm) one
        body = body[len(body) // 2:]
    for r in body:
        name = r[idx["Kernel Name"]].replace("dca::", "").replace("void ", "").split("(")[0]
        e = {"name": name, "t_us": g(r, "gpu__time_duration.sum") * tmul,
             "dram_read_MB": g(r, "dram__bytes_read.sum") * rmul / 1e6,
             "dram_write_MB": g(r, "dram__bytes_write.sum") * wmul / 1e6,
             "dram_pct": g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
             "tensor_pct": g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
             "sm_pct": g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
             "issue_pct": g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "inst_M": g(r, "smsp__inst_executed.sum") / 1e6,
             "l2_MB": g(r, "lts__t_bytes.sum") * mul.get(units[idx["lts__t_bytes.sum"]], 1.0) / 1e6 if "lts__t_bytes.sum" in idx else None,
             "regs": g(r, "launch__registers_per_thread")}
        out.append(e)
        traffic[name].append({"dram_read_bytes": e["dram_read_MB"] * 1e6, "dram_write_bytes": e["dram_write_MB"] * 1e6,
                              "time_us": e["t_us"], "tensor_pipe_pct": e["tensor_pct"], "dram_pct": e["dram_pct"]})
    json.dump(out, open(prefix + "_ncu_forward.json", "w"), indent=1)
    json.dump(traffic, open(prefix + "_ncu_traffic.json", "w"), indent=1)
    agg = collections.OrderedDict()
    for e in out:
        a = agg.setdefault(e["name"], dict(n=0, t=0.0, tensor=0.0, dram=0.0, rd=0.0, wr=0.0))
        a["n"] += 1; a["t"] += e["t_us"]; a["tensor"] += e["tensor_pct"] * e["t_us"]; a["dram"] += e["dram_pct"] * e["t_us"]
        a["rd"] += e["dram_read_MB"]; a["wr"] += e["dram_write_MB"]
    tot = sum(a["t"] for a in agg.values())
    print("| kernel | launches / forward | total us | share | tensor pipe % (time-weighted) | ncu DRAM % | DRAM MB read / written per launch |")
    print("|---|---|---|---|---|---|---|")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
        print(f"| `{k}` | {a['n']} | {a['t']:.0f} | {100 * a['t'] / tot:.1f}% | {a['tensor'] / a['t']:.1f} | "
              f"{a['dram'] / a['t']:.1f} | {a['rd'] / a['n']:.0f} / {a['wr'] / a['n']:.0f} |")
    print(f"\nTotal {tot / 1e3:.2f} ms under ncu (cold-cache, serialised).")


if __name__ == "__main__":
    main()
