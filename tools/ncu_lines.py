"""Aggregate an `ncu --page source --csv` SASS listing by CUDA source line.

    python tools/ncu_lines.py <ncu_source.csv> <nvdisasm --print-line-info listing of the same kernel> [top]

ncu's CSV export of the source page carries per-SASS-instruction counters but no source correlation; the
nvdisasm listing (built with -lineinfo) carries the line of every instruction.  The two are joined by the
instruction offset inside the kernel.
"""
import csv
import re
import sys
from collections import defaultdict


def main():
    ncu_csv, sass, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40
    rows = list(csv.reader(open(ncu_csv)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    base = int(body[0][col["Address"]], 16)
    line_of = {}
    cur = ("?", 0)
    for ln in open(sass):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if m:
            line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())
    agg = defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    ops = defaultdict(int)
    for r in body:
        off = int(r[col["Address"]], 16) - base
        key, text = line_of.get(off, (("?", 0), r[col["Source"]]))
        wi, ti, sm = int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]]), int(r[col["# Samples"]])
        for k, v in enumerate((wi, ti, sm)):
            agg[key][k] += v
            tot[k] += v
        ops[r[col["Source"]].split()[0 if not r[col["Source"]].strip().startswith("@") else 1].split(".")[0]] += wi
    print("total warp-inst %d thread-inst %d samples %d  (lanes/inst %.2f)" % (tot[0], tot[1], tot[2], tot[1] / max(tot[0], 1)))
    print("%-28s %8s %8s %8s %6s" % ("source line", "warp%", "thread%", "stall%", "lanes"))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s %8.2f %8.2f %8.2f %6.1f" % ("%s:%d" % key, 100 * v[0] / tot[0], 100 * v[1] / tot[1], 100 * v[2] / max(tot[2], 1), v[1] / max(v[0], 1)))
    print("\nopcode mix (warp-level):")
    for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:25]:
        print("  %-10s %6.2f%%" % (k, 100 * v / tot[0]))


if __name__ == "__main__":
    main()
