"""Bucket an `ncu --page source --csv` SASS listing of photon_kernel into the phases of the photon loop.

    python tools/ncu_regions.py <ncu_source.csv> <nvdisasm --print-line-info listing>

Lines are attributed through the innermost inlined location nvdisasm prints; device helpers (rng, rotate,
sqrt...) are attributed to the phase by the number of active lanes they run with.
"""
import csv
import re
import sys
from collections import defaultdict


def load(ncu_csv, sass):
    rows = list(csv.reader(open(ncu_csv)))
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    body = rows[2:]
    base = int(body[0][col["Address"]], 16)
    line_of = {}
    cur = ("?", 0)
    stack = []
    for ln in open(sass):
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            if "inlined at" in m.group(3):
                # innermost first; the following lines give the callers
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                stack = [cur]
            elif stack and ln.strip().startswith("//## File") and "inlined" not in ln and False:
                pass
            else:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                stack = [cur]
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    out = []
    for r in body:
        off = int(r[col["Address"]], 16) - base
        out.append((off, r[col["Source"]].strip(), int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]]),
                    int(r[col["# Samples"]]), line_of.get(off, ("?", 0))))
    return out


if __name__ == "__main__":
    data = load(sys.argv[1], sys.argv[2])
    tot = sum(d[2] for d in data)
    tt = sum(d[3] for d in data)
    by = defaultdict(lambda: [0, 0, 0, 0])
    for off, src, wi, ti, sm, key in data:
        lanes = ti / wi if wi else 0
        k = "%s:%d" % key
        by[k][0] += wi
        by[k][1] += ti
        by[k][2] += sm
        by[k][3] += 1
    print("static instr %d, warp-inst %d, lanes/inst %.2f" % (len(data), tot, tt / tot))
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
        print("%-28s n=%3d warp%% %6.2f lanes %5.1f  per-exec-instr %5.1f" % (k, v[3], 100 * v[0] / tot, v[1] / max(v[0], 1), v[0] / max(1, max(d[2] for d in data if "%s:%d" % d[5] == k))))
