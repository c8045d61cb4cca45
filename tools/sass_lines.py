"""Dev tool: attribute the SASS of one kernel to source lines (needs -lineinfo).
usage: python tools/sass_lines.py <cubin-name-fragment> <kernel-name-fragment> [per]
Splits the opcodes into the integer ALU pipe (the bound of the step kernels) and the rest."""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALU = {'LOP3', 'IADD3', 'SEL', 'ISETP', 'SHF', 'VIADD', 'VIMNMX', 'VIADDMNMX', 'PRMT', 'LEA', 'FSEL', 'VABSDIFF',
       'IABS', 'MOV', 'ISCADD', 'BMSK', 'SGXT', 'FLO', 'IMNMX', 'PLOP3', 'P2R', 'R2P', 'CS2R'}


def main():
    frag, kern = sys.argv[1], sys.argv[2]
    per = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "posggym_baselines_b200", "libposg_b200.so")],
                   cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if frag in f and f.count("-") == 0][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    secs = re.split(r"\n\s*\.section\s+\.text\.", txt)
    body = [s for s in secs if kern in s.split("\n", 1)[0]][0]
    cur, alu, tot = None, collections.Counter(), collections.Counter()
    for line in body.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            tot[cur] += 1
            if m.group(2) in ALU:
                alu[cur] += 1
    src = {}
    for k, v in sorted(tot.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        f, ln = k
        if f not in src:
            for root, _, files in os.walk(os.path.join(ROOT, "posggym_baselines_b200", "csrc")):
                if f in files:
                    src[f] = open(os.path.join(root, f)).read().split("\n")
        text = src.get(f, [""] * (ln + 1))[ln - 1].strip()[:90] if f in src else ""
        print(f"{f}:{ln:4d} alu {alu[k] / per:6.1f} total {v / per:6.1f}  {text}")
    print("sum alu", sum(alu.values()) / per, "total", sum(tot.values()) / per)


main()
