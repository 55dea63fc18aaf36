"""Static view of a kernel's SASS: loops (backward branches) with instruction counts and opcode mix.
usage: sass_loops.py <cuobjdump -sass output> <kernel name substring> [min_len]"""
import collections
import re
import sys

txt, kern = open(sys.argv[1], errors="ignore").read().split("\n"), sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ins, on = [], False
for l in txt:
    if "Function :" in l:
        on = kern in l
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f"{kern}: {len(ins)} instructions")
PIPE = {"FADD": "fma", "FMUL": "fma", "FFMA": "fma", "IMAD": "fma", "HFMA2": "fma", "IADD3": "alu", "IADD": "alu", "LOP3": "alu", "SHF": "alu", "PRMT": "alu",
        "FMNMX": "alu", "FSEL": "alu", "SEL": "alu", "ISETP": "alu", "FSETP": "alu", "VIMNMX": "alu", "VIADDMNMX": "alu", "VIADD": "alu", "LEA": "alu", "MOV": "alu", "PLOP3": "alu",
        "POPC": "xu", "FLO": "xu", "BREV": "xu", "LDS": "lsu", "STS": "lsu", "LDG": "lsu", "STG": "lsu", "SHFL": "lsu", "LDC": "lsu", "VOTE": "misc", "BRA": "cbu", "BSSY": "cbu", "BSYNC": "cbu",
        "WARPSYNC": "cbu", "R2P": "alu", "P2R": "alu", "FMNMX3": "alu", "UMOV": "u", "ULOP3": "u", "UIADD3": "u", "USHF": "u", "UISETP": "u", "R2UR": "u", "S2R": "misc", "NOP": "misc"}
for addr, t in ins:
    if "BRA" in t:
        m = re.search(r"0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            body = [x for a, x in ins if tgt <= a <= addr]
            if len(body) < min_len:
                continue
            ops = collections.Counter()
            pipes = collections.Counter()
            for x in body:
                op = re.sub(r"^@!?U?P\d+\s+", "", x).split()[0].split(".")[0]
                ops[op] += 1
                pipes[PIPE.get(op, "?" + op)] += 1
            print(f"loop {tgt:#x}..{addr:#x}: {len(body)} instr; pipes {dict(pipes)}")
            print("   ", ", ".join(f"{k}:{v}" for k, v in ops.most_common(40)))
