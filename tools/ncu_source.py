#!/usr/bin/env python
"""Per-region view of an ncu source page (SASS): instructions executed and stall samples between barriers.

    python tools/ncu_source.py report.ncu-rep 'regex:kernel' [launch_index]
Regions are delimited by BAR.SYNC / EXIT so that the phases of a kernel can be told apart without source lines.
"""
import csv
import io
import subprocess
import sys


def main(rep, kernel, skip="0"):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    lines = raw.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
    lines = lines[:end]
    rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))) if r.get("Instructions Executed") not in (None, "")]
    stall_cols = [c for c in rows[0] if c.startswith("stall_") and "Not Issued" not in c]
    region, regions = dict(n=0, inst=0, samples=0, stalls={}, ops={}), []
    total_inst = sum(int(r["Instructions Executed"]) for r in rows)
    total_samples = sum(int(r["# Samples"]) for r in rows)
    for r in rows:
        op = r["Source"].split()[0] if not r["Source"].strip().startswith("@") else r["Source"].split()[1]
        op = op.split(".")[0]
        region["n"] += 1
        ie = int(r["Instructions Executed"])
        region["inst"] += ie
        region["samples"] += int(r["# Samples"])
        region["ops"][op] = region["ops"].get(op, 0) + ie
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                region["stalls"][c] = region["stalls"].get(c, 0) + v
        if "BAR.SYNC" in r["Source"] or "EXIT" in r["Source"]:
            regions.append(region)
            region = dict(n=0, inst=0, samples=0, stalls={}, ops={})
    if region["n"]:
        regions.append(region)
    print(f"total warp-instructions {total_inst}, samples {total_samples}")
    for i, g in enumerate(regions):
        if g["inst"] == 0:
            continue
        top = sorted(g["stalls"].items(), key=lambda kv: -kv[1])[:4]
        ops = sorted(g["ops"].items(), key=lambda kv: -kv[1])[:6]
        print(f"region {i:2d}: sass {g['n']:5d}  inst {g['inst']:9d} ({g['inst'] / total_inst:5.1%})  samples {g['samples']:6d} "
              f"({g['samples'] / max(total_samples, 1):5.1%})  stalls {', '.join(f'{k[6:]}={v}' for k, v in top)}")
        print(f"           ops {', '.join(f'{k}={v}' for k, v in ops)}")


if __name__ == "__main__":
    main(*sys.argv[1:])
