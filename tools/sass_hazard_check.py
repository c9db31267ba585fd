#!/usr/bin/env python
"""Scan the SASS of libmde_b200.so for an under-stalled `CS2R Rd, SRZ` (64-bit register zeroing)
followed too closely by a reader of Rd / Rd+1.

Found on B200 with CUDA 12.9 ptxas (profiles/r01_ptxas_cs2r_hazard.md): ptxas scheduled a dependent
IADD3 five issue cycles after `CS2R R8, SRZ`; the hardware does not interlock fixed-latency RAW
hazards, the IADD3 read the stale register, and the kernel faulted (cudaErrorInvalidAddressSpace).
This script reports every site where a read follows within MIN_CYCLES so the source can be reshaped.
"""
import re
import subprocess
import sys

MIN_CYCLES = 8


def parse(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    fn, ins, funcs = None, [], {}
    lines = out.split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.search(r"Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            ins = funcs.setdefault(fn, [])
        m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", ln)
        if m and fn and i + 1 < len(lines):
            m2 = re.search(r"/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                ctrl = int(m2.group(1), 16) >> 41
                ins.append((int(m.group(1), 16), m.group(2).strip(), ctrl & 0xF))
                i += 2
                continue
        i += 1
    return funcs


def regs_read(text):
    """registers appearing as SOURCE operands (everything after the first operand; plus address regs)"""
    body = re.sub(r"^@!?U?P\d+\s+", "", text)
    parts = body.split(None, 1)
    if len(parts) < 2:
        return set()
    ops = parts[1]
    first, _, rest = ops.partition(",")
    srcs = rest
    # stores / reds read their first operand too (address / value)
    if re.match(r"(ST|RED|ATOM)", parts[0]):
        srcs = ops
    else:
        srcs += " " + " ".join(re.findall(r"\[(.*?)\]", first))
    out = set()
    for m in re.finditer(r"\bR(\d+)(\.64)?", srcs):
        r = int(m.group(1))
        out.add(r)
        if m.group(2):
            out.add(r + 1)
    return out


def main(path):
    bad = fatal = 0
    for fn, ins in parse(path).items():
        for k, (pc, text, stall) in enumerate(ins):
            m = re.match(r"CS2R R(\d+), SRZ", text)
            if not m:
                continue
            rd = int(m.group(1))
            cyc = stall
            for pc2, t2, s2 in ins[k + 1:k + 12]:
                if cyc >= MIN_CYCLES:
                    break
                rr = regs_read(t2)
                wide = ("WIDE" in t2 or ".64" in t2 or t2.startswith("D"))
                if rd in rr or (rd + 1) in rr or (wide and (rd - 1) in rr and False):
                    print("HAZARD %s\n   %05x: %s\n   %05x: %s   (%d cycles after)" % (fn[:90], pc, text, pc2, t2, cyc))
                    bad += 1
                    fatal += 1 if cyc < 6 else 0
                    break
                if re.match(r"(BRA|EXIT|RET|CALL|BSYNC)", re.sub(r"^@!?U?P\d+\s+", "", t2)):
                    break
                cyc += s2
    print("%d potential CS2R under-stall site(s) (threshold %d cycles; the observed failure was at 5)" % (bad, MIN_CYCLES))
    return 1 if fatal else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "pymde_b200/libmde_b200.so"))
