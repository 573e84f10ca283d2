"""Condenses an ncu launch list (--csv, --metrics gpu__time_duration.sum[,dram bytes,...]) into per-kernel totals."""
import csv
import sys
from collections import OrderedDict


def main(path, out=None):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    iname, imetric, ival = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    iid = hdr.index("ID")
    per = OrderedDict()
    for r in rows[1:]:
        if len(r) <= ival:
            continue
        k = r[iname]
        d = per.setdefault(k, {"ids": set()})
        d["ids"].add(r[iid])
        try:
            d[r[imetric]] = d.get(r[imetric], 0.0) + float(r[ival].replace(",", ""))
        except ValueError:
            pass
    lines = ["kernel,launches,total_us,avg_us,dram_read_MB_per_launch,dram_write_MB_per_launch"]
    tot = 0.0
    for k, d in sorted(per.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
        n = len(d["ids"])
        t = d.get("gpu__time_duration.sum", 0.0) / 1e3  # ns -> us
        tot += t
        rd = d.get("dram__bytes_read.sum", 0.0) / n / 1e6
        wr = d.get("dram__bytes_write.sum", 0.0) / n / 1e6
        lines.append(f'"{k[:90]}",{n},{t:.1f},{t / n:.2f},{rd:.2f},{wr:.2f}')
    lines.append(f'"TOTAL",,{tot:.1f},,,')
    txt = "\n".join(lines)
    if out:
        open(out, "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main(*sys.argv[1:])
