"""Where the issue slots of one photon_kernel capture go, by the number of lanes the instructions ran with.

    ncu -i <report.ncu-rep> --page source --csv > src.csv
    python tools/ncu_lanes.py src.csv [title]

Buckets follow profiles/r2_v9_loop_breakdown.txt: >= 28 lanes (segment block), 22-28 (scattering / queue refill),
15-22 (reduction, queue pop), 4-15, < 4 (tail block: boundary, retire, launch).  The number of warp-iterations of the
photon loop is taken from the most-executed backward branch.
"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
kernel = rows[0][1]
col = {h: i for i, h in enumerate(rows[1])}
body = rows[2:]
base = int(body[0][col["Address"]], 16)
inst = []
for r in body:
    src = r[col["Source"]].strip()
    tok = src.split()
    op = (tok[1] if tok[0].startswith("@") else tok[0]).split(".")[0]
    inst.append((int(r[col["Address"]], 16) - base, src, op, int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]])))
tot_w = sum(i[3] for i in inst)
tot_t = sum(i[4] for i in inst)
# loop trips: the most-executed backward BRA (target below its own address)
trips = 0
for off, src, op, w, t in inst:
    if op == "BRA" and "0x" in src:
        tgt = int(src.split("0x")[-1].split()[0].rstrip(";"), 16) - base
        if tgt <= off:
            trips = max(trips, w)
buckets = [("segment, deposit arithmetic, queue test (>= 28 lanes)", 28, 33), ("scattering block / queue refill (22-28 lanes)", 22, 28),
           ("reduction + queue pop (15-22 lanes)", 15, 22), ("4-15 lanes", 4, 15), ("tail block: boundary / retire / launch (< 4 lanes)", 0, 4)]
print("%s\n%s" % (sys.argv[2] if len(sys.argv) > 2 else "", kernel))
print("static SASS instructions %d, warp instructions %.4e, thread instructions %.4e, lanes per instruction %.2f" % (len(inst), tot_w, tot_t, tot_t / tot_w))
print("most-executed backward branch: %.4e warp-trips (one trip of the shipped common kernels = two loop iterations, `#pragma unroll 2`)\n" % trips)
iters = 2.0 * trips
print("=> about %.4e warp-iterations, %.1f warp instructions per iteration\n" % (iters, tot_w / max(iters, 1)))
print("%-58s %10s %8s %7s %10s" % ("phase (by active lanes)", "warp-inst", "share", "lanes", "per iter"))
for name, lo, hi in buckets:
    sel = [i for i in inst if i[3] and lo <= i[4] / i[3] < hi]
    w, t = sum(i[3] for i in sel), sum(i[4] for i in sel)
    print("%-58s %10.3e %7.1f%% %7.1f %10.1f" % (name, w, 100.0 * w / tot_w, t / max(w, 1), w / max(iters, 1)))
ops = defaultdict(int)
for i in inst:
    ops[i[2]] += i[3]
print("\nopcode      share of warp instructions")
for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:22]:
    print("%-10s %6.2f%%" % (k, 100.0 * v / tot_w))
