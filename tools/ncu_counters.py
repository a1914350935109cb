"""Write the per-photon counters of one `ncu --set full` capture of the photon kernel into profiles/kernel_counters.json.

    python tools/ncu_counters.py <report.ncu-rep> <workload> <photons in the captured launch> <label of the capture>
"""
import csv
import json
import os
import subprocess
import sys

rep, workload, nph, label = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}


def val(name):
    v, u = float(vals[col[name]].replace(",", "")), units[col[name]]
    return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)


winst = val("smsp__inst_executed.sum")
tinst = winst * val("smsp__thread_inst_executed_per_inst_executed.ratio")
dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "kernel_counters.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[workload] = dict(warp_inst_per_photon=winst / nph, thread_inst_per_photon=tinst / nph, lanes_per_warp_inst=tinst / winst,
                   dram_bytes_per_photon=dram / nph, issue_active_pct=val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   kernel=vals[col["Kernel Name"]], photons=nph, source=label)
json.dump(d, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(d[workload], indent=1))
