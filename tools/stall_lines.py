"""Dev tool: attribute the warp-stall samples of an ncu capture to source lines.
usage: python tools/stall_lines.py <ncu source-page csv (sass)> <cubin-name-fragment> <kernel-name-fragment> [top]
The csv comes from `ncu -i X.ncu-rep --page source --csv`; the library in the tree must be the build that was
captured (the SASS of the kernel is aligned instruction by instruction)."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, frag, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = {n: i for i, n in enumerate(rows[hdr])}
    prof = [(r[cols["Source"]].split()[0 if not r[cols["Source"]].lstrip().startswith("@") else 1], int(r[cols["# Samples"]] or 0),
             int(r[cols["Instructions Executed"]] or 0), int(r[cols["stall_long_sb"]] or 0))
            for r in rows[hdr + 1:] if r and r[0].startswith("0x")]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "posggym_baselines_b200", "libposg_b200.so")],
                   cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if frag in f][0]
    txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    secs = re.split(r"\n\s*\.section\s+\.text\.", txt)
    body = [s for s in secs if kern in s.split("\n", 1)[0]][0]
    cur, sass = None, []
    for line in body.split("\n"):
        m = re.search(r'//## File "([^"]+)", line (\d+)', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            sass.append((cur, m.group(2)))
    print("instructions: profile", len(prof), "library", len(sass))
    n = min(len(prof), len(sass))
    bad = sum(1 for i in range(n) if prof[i][0].split(".")[0] != sass[i][1].split(".")[0])
    print("opcode mismatches", bad)
    samples, execd, lsb = collections.Counter(), collections.Counter(), collections.Counter()
    for i in range(n):
        samples[sass[i][0]] += prof[i][1]; execd[sass[i][0]] += prof[i][2]; lsb[sass[i][0]] += prof[i][3]
    total = sum(samples.values())
    src = {}
    for (f, ln), v in samples.most_common(top):
        if f not in src:
            for root, _, files in os.walk(os.path.join(ROOT, "posggym_baselines_b200", "csrc")):
                if f in files:
                    src[f] = open(os.path.join(root, f)).read().split("\n")
        text = src[f][ln - 1].strip()[:100] if f in src else ""
        print(f"{100 * v / total:5.1f}%  long_sb {lsb[(f, ln)]:6d}  exec {execd[(f, ln)]:8d}  {f}:{ln}  {text}")
    byfile = collections.Counter()
    for (f, ln), v in samples.items():
        byfile[f] += v
    print({f: round(100 * v / total, 1) for f, v in byfile.items()})


main()
