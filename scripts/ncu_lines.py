"""Per-source-line instruction / stall-sample totals for one kernel of an .ncu-rep.

`ncu --page source --csv` only exports the SASS view; this joins it (by instruction offset)
with `nvdisasm -g` line info of the object file that holds the kernel.

  python scripts/ncu_lines.py <report.ncu-rep> <object.o> <kernel-substring> [top] [mangled-substring]

(the mangled substring selects the SASS section when the object holds several template
instantiations, e.g. k_astarILj32E)
"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
mangled = sys.argv[5] if len(sys.argv) > 5 else kern

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout

# offset -> (file, line) for the wanted function
line_of = {}
in_fn = False
cur = None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        in_fn = mangled in m.group(1)
        continue
    if not in_fn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the report may hold several kernels: pick the block whose "Kernel Name" matches
blocks, cur_rows, name = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        if cur_rows:
            blocks.append((name, cur_rows))
        name, cur_rows = r[1], []
    elif cur_rows is not None:
        cur_rows.append(r)
if cur_rows:
    blocks.append((name, cur_rows))
name, blk = [b for b in blocks if kern in b[0]][0]
h = blk[0]
iA, iS = h.index("Address"), h.index("Source")
iI, iT, iN = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
base = int(blk[1][iA], 16)
agg = {}
tot_i = tot_s = 0
for r in blk[1:]:
    off = int(r[iA], 16) - base
    key = line_of.get(off, ((None, 0), ""))[0]
    a = agg.setdefault(key, {"inst": 0, "thr": 0, "samp": 0, "stalls": {}})
    a["inst"] += int(r[iI])
    a["thr"] += int(r[iT])
    a["samp"] += int(r[iN])
    for i, c in stall_cols:
        v = int(r[i]) if r[i].isdigit() else 0
        if v:
            a["stalls"][c] = a["stalls"].get(c, 0) + v
    tot_i += int(r[iI])
    tot_s += int(r[iN])

src_cache = {}


def src(key):
    if not key or not key[0]:
        return "?"
    f = key[0]
    if f not in src_cache:
        cands = glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), f))
        src_cache[f] = open(cands[0]).read().splitlines() if cands else []
    L = src_cache[f]
    return L[key[1] - 1].strip()[:90] if 0 < key[1] <= len(L) else ""


print(f"kernel {name[:60]}: {tot_i} warp-instructions, {tot_s} stall samples")
print("  inst%  samp%  lanes  top stalls | line: source")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samp"])[:top]:
    st = sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3]
    sts = " ".join(f"{c[6:]}:{100 * v / max(a['samp'], 1):.0f}" for c, v in st)
    print(f"{100 * a['inst'] / tot_i:6.1f} {100 * a['samp'] / max(tot_s, 1):6.1f} {a['thr'] / max(a['inst'], 1):6.1f}  {sts:36s} | "
          f"{key[0] if key else '?'}:{key[1] if key else 0}: {src(key)}")
