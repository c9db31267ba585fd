#!/usr/bin/env python
"""Turn an `ncu --set full` report into the committed artefacts bench.py and the judge read:

    python tools/ncu_extract.py gpurun_out/r02_pull_c2.ncu-rep C2 profiles/r02_ncu_pull_c2_raw.csv

writes the raw page as CSV (argument 3) and adds / replaces entry `C2` of profiles/r02_kernel_profile.json:
kernel name, ncu duration, dram bytes read / written (`dram__bytes_read.sum + dram__bytes_write.sum`: the
`roofline.traffic` of the bench line), registers, issue-active, top stall reasons.  Runs here (no GPU needed)."""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def to_bytes(value, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value.replace(",", "")) * mult.get(unit, 1)


def main():
    rep, key, raw_out = sys.argv[1], sys.argv[2], sys.argv[3]
    if rep.endswith(".csv"):  # an already exported raw page (e.g. the r01 captures whose .ncu-rep was not kept)
        txt = open(rep).read()
        raw_out = rep
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(raw_out, "w").write(txt)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}

    def get(name, conv=float):
        i = col.get(name)
        if i is None or vals[i] == "":
            return None
        return conv(vals[i].replace(",", ""))

    rd = to_bytes(vals[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(vals[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    dur = get("gpu__time_duration.sum")
    dur_unit = units[col["gpu__time_duration.sum"]]
    dur_us = dur * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(dur_unit, 1.0)
    stalls = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(vals[i].replace(",", ""))
              for h, i in col.items() if "smsp__average_warps_issue_stalled_" in h and h.endswith("per_issue_active.ratio")
              and vals[i] not in ("", "n/a")}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:5]
    entry = {"kernel": vals[col["Kernel Name"]], "grid": vals[col["Grid Size"]], "block": vals[col["Block Size"]],
             "ncu_duration_us": dur_us, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": rd + wr,
             "registers_per_thread": get("launch__registers_per_thread"),
             "warp_instructions": get("smsp__inst_executed.sum"),
             "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "sm_cycles_active_avg": get("sm__cycles_active.avg"), "sm_cycles_elapsed_max": get("sm__cycles_elapsed.max"),
             "lts_sectors": get("lts__t_sectors.sum"),
             "stall_warps_per_issue_top5": top,
             "source": os.path.relpath(raw_out, REPO), "report": os.path.basename(rep),
             "how": "ncu --set full --clock-control none --import-source on (one launch, cold caches under ncu replay)"}
    path = os.path.join(REPO, "profiles", "r02_kernel_profile.json")
    prof = json.load(open(path)) if os.path.exists(path) else {}
    prof[key] = entry
    json.dump(prof, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
