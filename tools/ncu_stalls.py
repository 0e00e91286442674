#!/usr/bin/env python
"""Top warp-stall reasons per profiled launch of an `ncu --set full` report (stalled warps per issue-active cycle)."""
import csv
import io
import subprocess
import sys


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        vals = sorted(((float(r[idx[c]] or 0), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")])
                       for c in cols), reverse=True)[:6]
        print(f"{r[idx['Kernel Name']][:44]:44s} " + "  ".join(f"{n}={v:.2f}" for v, n in vals))


if __name__ == "__main__":
    main(sys.argv[1])
