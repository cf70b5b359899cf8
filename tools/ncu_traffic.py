#!/usr/bin/env python
"""Regenerates profiles/ncu_traffic.json (measured DRAM bytes per unit of work of each kernel) from ncu reports.
   python tools/ncu_traffic.py gpurun_out/r1h.ncu-rep:59520000 gpurun_out/r1h_solver.ncu-rep:8664016
   (report:units, units = 3-d cells of the tx_sample launch resp. 2-d points of the padded tx0.1v3 block)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"tracer_fast_kernel": "TRACER_UPDATE", "tracer_column_kernel": "TRACER_UPDATE", "momentum_column_kernel": "MOMENTUM_COLUMN",
         "impvmixt_kernel<0": "VMIX_TRACER_IMPLICIT", "impvmixt_kernel<(bool)0": "VMIX_TRACER_IMPLICIT",
         "momentum_finish_kernel": "MOMENTUM_FINISH", "state_3d_kernel": "STATE",
         "pcsi_iter2_kernel<0>": "PCSI_PASS2_KERNEL", "pcsi_iter2_kernel<(bool)0>": "PCSI_PASS2_KERNEL"}
out = {"_source": "ncu --set full: (dram__bytes_read.sum + dram__bytes_write.sum) / units of the captured launch; tools/ncu_traffic.py " + " ".join(sys.argv[1:])}
for arg in sys.argv[1:]:
    rep, units = arg.rsplit(":", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt))); hdr, un = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    def val(r, m):
        v = float(r[idx[m]].replace(",", "")); u = un[idx[m]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
    for r in rows[2:]:
        kn = r[idx["Kernel Name"]].replace("void ", "").split("(")[0]
        key = next((v for k, v in NAMES.items() if kn.startswith(k)), None)
        if key and key not in out:
            b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
            out[key] = {"bytes_per_unit": round(b / float(units), 2), "unit": "2-d point" if "PCSI" in key else "3-d cell",
                        "report": os.path.basename(rep), "kernel": kn}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
