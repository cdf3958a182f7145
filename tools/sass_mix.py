#!/usr/bin/env python3
"""Instruction-mix summary of an address range of a SASS dump (development aid).
usage: sass_mix.py file.sass 0xSTART 0xEND"""
import re, sys, collections
ALU = ("LOP3", "PRMT", "SHF", "VIADDMNMX", "VIMNMX3", "IADD3", "LEA", "ISETP", "SEL", "PLOP3", "IABS", "BREV", "FLO", "POPC")
ALU_Q = ("VIMNMX", "HMNMX2")          # quarter-cycle ops on the ALU pipe (2-input)
FMA = ("IMAD", "VIADD", "HADD2", "HFMA2", "FFMA", "FADD", "FMUL")
LSU = ("LDS", "STS", "LDG", "STG", "LDC", "ATOM", "RED", "LDL", "STL")
def main():
    path, lo, hi = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16)
    cnt = collections.Counter(); pipe = collections.Counter()
    for line in open(path):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m: continue
        a = int(m.group(1), 16)
        if not (lo <= a < hi): continue
        op = m.group(3); base = op.split(".")[0]
        cnt[op.split(".")[0] + ("." + op.split(".")[1] if base in ("VIMNMX", "VIADD", "IMAD") and "." in op else "")] += 1
        if base == "VIMNMX" or base == "HMNMX2": pipe["alu_quarter"] += 1
        elif base == "VIADDMNMX" or base == "VIMNMX3" or base in ALU: pipe["alu"] += 1
        elif base in FMA: pipe["fma"] += 1
        elif base in LSU: pipe["lsu"] += 1
        else: pipe["other"] += 1
    tot = sum(cnt.values())
    print("instructions:", tot, dict(pipe))
    print("ALU-pipe cycles/warp (0.5*alu + 0.25*quarter): %.1f   FMA: %.1f" % (0.5 * pipe["alu"] + 0.25 * pipe["alu_quarter"], 0.5 * pipe["fma"]))
    for k, v in cnt.most_common(25): print("  %-16s %d" % (k, v))
main()
