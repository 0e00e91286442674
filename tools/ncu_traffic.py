#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report: DRAM bytes per launch (dram__bytes_read.sum +
dram__bytes_write.sum, median over the captured launches) of every step kernel, keyed by bench.py's phase names.

    python tools/ncu_traffic.py gpurun_out/prof_t_lin_r02.ncu-rep t_lin [profiles/traffic.json]
"""
import csv
import io
import json
import os
import statistics
import subprocess
import sys

PHASE_OF = [("k_mc_", None), ("k_acyclic", "acyclic"), ("k_pair_dist", "pair_dist"), ("k_pair_finish", "pair_kernel"), ("k_phi", "phi_update"),
            ("k_prologue", "scores"), ("k_edge_probs", "scores")]


def main(rep, workload, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def to_bytes(v, unit):
        v = float(v.replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    per = {}
    mc_seen = 0
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        b = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + \
            to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        for pat, ph in PHASE_OF:
            if name.startswith(pat) or ("::" + pat) in name or pat in name:
                if ph is None:
                    # Monte-Carlo pass kernels: the MODE template argument (2nd) tells the Theta pass (0) from the Z pass
                    import re
                    mm = re.search(r"k_mc_\w+<(?:\(int\))?\d+, (?:\(int\))?(\d)", name)
                    mode = int(mm.group(1)) if mm else 1
                    ph = "mc_theta" if mode == 0 else "mc_z"
                    mc_seen += 1
                per.setdefault(ph, []).append(b)
                break
    table = {}
    if os.path.exists(out):
        table = json.load(open(out))
    table["_note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per launch (median over the captured launches) from "
                      "`ncu --set full` captures of `python bench.py --workload W --steps 6 --warmup 3`; one GPU only")
    table[workload] = {k: int(statistics.median(v)) for k, v in per.items()}
    table[workload]["_source"] = os.path.basename(rep)
    json.dump(table, open(out, "w"), indent=1)
    print(json.dumps(table[workload], indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else
         os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"))
