"""Instruction families of every kernel of this library, from the objects the in-tree build leaves under
binocular3dgs_b200/csrc/build/ (cuobjdump + cu++filt; no GPU needed).
usage: python tools/sass_summary.py > profiles/r02_own_kernels_sass_summary.txt
       python tools/sass_summary.py --full composite > profiles/r02_composite_sm100a.sass"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "binocular3dgs_b200", "csrc", "build")
# opcodes reported with their width / space modifiers; everything else by its bare mnemonic
KEEP_MODIFIERS = ("LDG", "STG", "LDS", "STS", "RED", "ATOM", "LDGMC", "UBLKCP", "SYNCS", "MEMBAR", "LD", "ST")
INTERESTING_FIRST = ("UBLKCP", "SYNCS", "LDGMC", "ACQBULK", "FFMA2", "FMUL2", "FADD2", "UTMA", "MATCH", "REDUX")


def demangle(names):
    if not names:
        return []
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, stdin=subprocess.DEVNULL)
    return out.stdout.splitlines() if out.returncode == 0 else names


def kernels_of(obj):
    text = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    cur, table = None, collections.OrderedDict()
    for line in text.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            table[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur is not None:
            table[cur].append(m.group(1).strip())
    return table


def family(instr):
    tok = instr.split()
    op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
    parts = op.split(".")
    if parts[0] in KEEP_MODIFIERS:
        keep = [p for p in parts[1:] if p in ("E", "64", "128", "U8", "U16", "CONSTANT", "STRONG", "ADD", "F32", "MIN", "MAX")]
        return ".".join([parts[0]] + keep)
    return parts[0]


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--full":
        obj = os.path.join(BUILD, sys.argv[2] + ".o")
        sys.stdout.write(subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout)
        return
    print("SASS of this library's kernels (sm_100a, nvcc 12.9, -O3 -lineinfo): instruction families per kernel, from the")
    print("objects of the in-tree build (tools/sass_summary.py).  Full listing of the composite kernels: r02_composite_sm100a.sass.")
    print("UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier, LDGMC = multimem.ld_reduce (NVLS),")
    print("ACQBULK = griddepcontrol.wait (programmatic dependent launch), FFMA2 / FMUL2 / FADD2 = packed FP32 (fma.rn.f32x2 ...).")
    print()
    for obj in sorted(glob.glob(os.path.join(BUILD, "*.o"))):
        table = kernels_of(obj)
        names = list(table)
        for mangled, pretty in zip(names, demangle(names)):
            instrs = [i for i in table[mangled] if not i.startswith("NOP")]
            if not instrs:
                continue
            counts = collections.Counter(family(i) for i in instrs)
            first = [k for k in INTERESTING_FIRST if k in counts]
            rest = [k for k, _ in counts.most_common() if k not in first and k not in ("BRA", "MOV", "IMAD", "IADD3", "ISETP", "LOP3", "BSSY", "BSYNC", "EXIT", "S2R", "LDC", "LDCU", "SHF", "LEA", "UMOV", "S2UR", "VIADD", "SEL", "PLOP3", "IABS", "R2UR", "UIADD3", "ULDC", "ULEA", "USHF", "ULOP3", "UISETP", "UIMAD", "CS2R", "PRMT", "CALL", "RET", "WARPSYNC", "BREAK", "BMOV", "I2FP", "F2I", "I2F", "UFLO", "FLO", "POPC", "BREV", "IMNMX", "VIMNMX", "UPLOP3", "USEL", "UPRMT", "LEPC", "NANOSLEEP", "YIELD", "ERRBAR", "CCTL", "DEPBAR", "P2R", "R2P", "UP2UR", "UR2UP", "FSEL", "FSETP", "FMNMX", "FCHK", "HFMA2", "UIADD3.64", "IADD")][:22]
            print("%s  [%s]" % (pretty, os.path.basename(obj)))
            print("    %d instructions: %s" % (len(instrs), ", ".join("%s %d" % (k, counts[k]) for k in first + rest)))
            print()


if __name__ == "__main__":
    main()
