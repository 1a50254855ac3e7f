"""Instruction mix of the hot loops of a kernel in an object file (development aid, runs without a GPU).

usage: python tools/sass_loop.py worm_algorithm_b200/csrc/worm_pack.o 'worm_pack_kernelILi0ELb0ELb0ELi256' [min_len]

Finds every backward branch (a loop), prints its length in SASS instructions and how they split over the
issue pipes: fma (IMAD*, FFMA ...), alu (LOP3, SHF, IADD3, ISETP, SEL, PRMT ...), lsu (LDS/STS/LDG/RED ...),
uniform datapath (U*), control.  `-v` as a last argument prints the loop body."""
import re
import subprocess
import sys


def pipe(op):
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UNPACK",):
        return "uni"
    if base in ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL", "IDP", "VIADD") or op.startswith("IMAD"):
        # VIADD is issued on either pipe by the hardware's choice; counted with fma here (ptxas uses it as a balancer)
        return "fma"
    if base in ("LDS", "STS", "LDG", "STG", "RED", "ATOM", "ATOMS", "ATOMG", "LD", "ST", "LDSM", "LDC", "LDCU"):
        return "lsu"
    if base in ("BRA", "BSSY", "BSYNC", "EXIT", "BAR", "WARPSYNC", "NOP", "CALL", "RET", "BREAK", "YIELD"):
        return "ctl"
    if base in ("MUFU", "I2F", "F2I", "I2FP", "F2FP", "POPC", "FLO", "BREV"):
        return "xu"
    if base in ("S2R", "S2UR", "CS2R", "SHFL", "VOTE", "MATCH", "R2UR", "REDUX"):
        return "misc"
    return "alu"


def main():
    obj, pat = sys.argv[1], sys.argv[2]
    verbose = sys.argv[-1] == "-v"
    min_len = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 40
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)
    for f in funcs[1:]:
        name = f.split("\n", 1)[0]
        if pat not in name:
            continue
        ins = []
        for line in f.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                text = m.group(2).strip()
                t = text.split()
                op = t[1] if t[0].startswith("@") else t[0]
                ins.append((int(m.group(1), 16), op, text))
        addr_index = {a: i for i, (a, _, _) in enumerate(ins)}
        print(name[:150], "-- %d instructions" % len(ins))
        for i, (a, op, text) in enumerate(ins):
            if op.startswith("BRA"):
                m = re.search(r"0x([0-9a-f]+)", text)
                if not m:
                    continue
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr_index:
                    j = addr_index[tgt]
                    body = ins[j:i + 1]
                    if len(body) < min_len:
                        continue
                    mix = {}
                    for _, o, _ in body:
                        mix[pipe(o)] = mix.get(pipe(o), 0) + 1
                    nlds = sum(1 for _, o, _ in body if o.startswith("LDS"))
                    print("  loop %#06x..%#06x: %4d instr  %s  (LDS %d)" % (tgt, a, len(body),
                          " ".join("%s=%d" % kv for kv in sorted(mix.items())), nlds))
                    if verbose:
                        for aa, _, tt in body:
                            print("      /*%04x*/ %s" % (aa, tt))


if __name__ == "__main__":
    main()
