#!/usr/bin/env python
"""Lists the loops of a kernel's SASS with their instruction mix and a modelled issue cost per iteration (B200 issue rates
measured by tools/microbench: FP32 / IADD 1 cycle, ALU-pipe integer / select ops 2, shared-memory loads on their own pipe).
    python tools/sass_loops.py <object or cubin> <kernel-name-substring> [min_instr] [dump_loop_index]"""
import collections
import re
import subprocess
import sys

COST2 = ("LOP3", "SHF", "SEL", "IMAD", "ISETP", "PRMT", "LEA", "SGXT", "I2F", "I2FP", "FSETP", "PLOP3", "VIMNMX", "IMNMX", "MOV", "FSEL", "P2R", "R2P", "BMSK", "POPC", "FLO", "FMNMX", "VIADD", "FCHK")
FP1 = ("FFMA", "FMUL", "FADD", "IADD3", "IADD", "HFMA2", "FMNMX3")


def cost(op):
    base = op.split(".")[0]
    if base in ("LDS", "LDSM"):
        return 0.0          # own pipe (2 cycles each there)
    if base in ("FFMA", "FMUL", "FADD", "IADD3", "FMNMX3"):
        return 1.0
    if base == "VIADD":
        return 1.0
    if base == "FSEL":
        return 1.4
    if base in ("F2I", "MUFU"):
        return 8.0
    if base in ("STS", "SHFL"):
        return 4.0
    if base in ("BRA", "BSSY", "BSYNC", "CALL", "RET", "NOP", "WARPSYNC"):
        return 1.0
    if base in ("LDG", "STG", "LDL", "STL", "LDC", "LDCU", "LD", "ST"):
        return 2.0
    return 2.0


def main():
    obj, name = sys.argv[1], sys.argv[2]
    min_n = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    dump = int(sys.argv[4]) if len(sys.argv) > 4 else -1
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s+Function : ", txt)[1:]
    for f in funcs:
        fname = f.split("\n")[0]
        if name not in fname:
            continue
        ins = []
        for l in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        addr = {a: i for i, (a, _) in enumerate(ins)}
        print(fname[:100], len(ins), "instructions")
        k = 0
        for i, (a, t) in enumerate(ins):
            m = re.search(r"BRA.*0x([0-9a-f]+)", t)
            if not m:
                continue
            tg = int(m.group(1), 16)
            if tg < a and tg in addr and i - addr[tg] + 1 >= min_n and i - addr[tg] + 1 < 2000:
                body = ins[addr[tg]:i + 1]
                c = collections.Counter()
                tot = 0.0
                lds = 0
                for _, tt in body:
                    op = tt.split()
                    o = op[1] if op[0].startswith("@") else op[0]
                    c[o.split(".")[0]] += 1
                    tot += cost(o)
                    lds += o.startswith("LDS")
                top = ", ".join(f"{k2} {v}" for k2, v in c.most_common(14))
                print(f"  loop {k}: {hex(tg)}-{hex(a)} {len(body)} instr, model cost {tot:.0f} cycles (+{lds} LDS on the LSU pipe): {top}")
                if k == dump:
                    for aa, tt in body:
                        print("     ", hex(aa), tt)
                k += 1


if __name__ == "__main__":
    main()
