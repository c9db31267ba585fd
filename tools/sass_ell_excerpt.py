#!/usr/bin/env python
"""Regenerate profiles/r02_sass_ell.md: excerpts of `cuobjdump -sass` for the ELL pull kernel
(distortion_ell_kernel<M=2, LOG1P, LOG, fast>) from the built object pymde_b200/csrc/_build/mde_ell.o.
Runs here (no GPU).  Landmarks are searched, not hard-coded, so it survives recompiles."""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(REPO, "pymde_b200", "csrc", "_build", "mde_ell.o")
OUT = os.path.join(REPO, "profiles", "r02_sass_ell.md")
TAG = "distortion_ell_kernelILi2ELi7ELi8ELb1E"


def kernel_lines():
    txt = subprocess.run(["cuobjdump", "-sass", OBJ], capture_output=True, text=True, check=True).stdout
    out, on = [], False
    for l in txt.splitlines():
        if "Function :" in l:
            on = TAG in l
            continue
        if on and re.match(r"^\s+/\*[0-9a-f]{4}\*/", l):
            out.append(re.sub(r"/\* 0x[0-9a-f]+ \*/", "", l).strip())
    return out


def main():
    L = kernel_lines()
    if not L:
        sys.exit("kernel not found in " + OBJ)
    find = lambda pat, start=0: next(i for i in range(start, len(L)) if re.search(pat, L[i]))
    doc = ["# r02 -- SASS of the ELL pull kernel (`distortion_ell_kernel<M=2, LOG1P, LOG, fast>`), sm_100a, nvcc 12.9\n",
           "`cuobjdump -sass pymde_b200/csrc/_build/mde_ell.o` (encodings dropped; regenerate with `tools/sass_ell_excerpt.py`).\n"
           "What to look for: the record stream and the X tile arrive by `UBLKCP.S.G` (cp.async.bulk, TMA) signalled on\n"
           "`SYNCS.ARRIVE.TRANS64` and waited with `SYNCS.PHASECHK.TRANS64.TRYWAIT`; inside the entry loop every memory access is an\n"
           "`LDS` (two record words per PAIR of entries, one neighbour-row gather per entry) -- no `LDG`, no `REDG`, no `ATOMS`; the\n"
           "only global accesses of a lane-slot are one `LDG.E.64.CONSTANT` (owner row) before the loop and one `REDG.E.ADD.F32x2`\n"
           "after it.\n"]
    # prologue
    i = find(r"LDG\.E\.128\.CONSTANT")
    j = find(r"UBLKCP", i)
    doc += ["## Prologue: ONE descriptor load, mbarrier init, the tile copy, then the first record copies\n", "```"]
    doc += L[max(0, i - 2):i + 2] + ["..."] + L[find(r"SYNCS\.EXCH", i) - 1:find(r"SYNCS\.EXCH", i) + 2] + ["..."] + L[j - 8:j + 2]
    k = find(r"@!UP0 UBLKCP|UBLKCP.*desc", j + 1)
    doc += ["..."] + L[k - 16:k + 2] + ["```\n"]
    # per record
    w = [i for i, l in enumerate(L) if "SYNCS.PHASECHK.TRANS64.TRYWAIT" in l and "[R" in l]
    w0 = w[0]
    doc += ["## Per record: wait for the slot; per lane-slot row: owner word from shared memory, owner row from global (L1)\n", "```"]
    doc += L[w0 - 4:w0 + 24] + ["```\n"]
    # entry loop: the back-edge whose body holds 4 MUFU.LG2 after the first wait
    loops = []
    for i, l in enumerate(L):
        m = re.search(r"@P\d BRA (0x[0-9a-f]+)", l)
        if m and i > w0:
            tgt = int(m.group(1), 16)
            j = next((q for q in range(i, -1, -1) if L[q].startswith("/*%04x*/" % tgt)), None)
            if j is not None and j < i and sum("MUFU.LG2" in x for x in L[j:i]) == 4 and not any("EX2" in x for x in L[j:i]):
                loops.append((j, i))
    if loops:
        j, i = loops[0]
        doc += ["## Entry loop, attractive class, 4 entries per trip: %d instructions = %.1f per directed entry, 4 MUFU each\n" % (
            i - j + 1, (i - j + 1) / 4.0), "```"] + L[j:i + 1] + ["```\n"]
    r = find(r"REDG", w0)
    u = find(r"UBLKCP", r)
    doc += ["## After the loop: class constant, ONE vector red per lane-slot; after the last row the slot is refilled\n", "```"]
    doc += L[r - 10:r + 2] + ["..."] + L[u - 14:u + 3] + ["```"]
    open(OUT, "w").write("\n".join(doc) + "\n")
    print("wrote", OUT, len(doc), "lines")


if __name__ == "__main__":
    main()
