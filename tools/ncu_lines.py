"""Aggregate an ncu source page (cuda,sass) dump by source line: executed warp-instructions and stall samples.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > both.csv; python tools/ncu_lines.py both.csv [N]"""
import collections
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except Exception:  # noqa: BLE001
        return 0


def main(path, top=45):
    cur, hdr, agg = None, None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ie, isamp = r.index("Instructions Executed"), r.index("# Samples")
            continue
        if r[0] != "" and hdr:
            try:
                ln = int(r[0])
            except ValueError:
                continue
            key = (cur, ln, r[1][:100])
            a = agg.get(key, (0, 0))
            agg[key] = (a[0] + num(r[ie]), a[1] + num(r[isamp]))
    tot = sum(v[0] for v in agg.values())
    ts = sum(v[1] for v in agg.values())
    print("total warp-instructions", tot, "samples", ts)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0]:9d} {100*v[0]/max(tot,1):5.1f}% samp={100*v[1]/max(ts,1):5.1f}%  {k[0]}:{k[1]}  {k[2]}")
    print("---- by stall samples")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:15]:
        print(f"{v[0]:9d} {100*v[0]/max(tot,1):5.1f}% samp={100*v[1]/max(ts,1):5.1f}%  {k[0]}:{k[1]}  {k[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
