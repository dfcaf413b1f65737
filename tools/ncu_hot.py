#!/usr/bin/env python
"""Attribute an ncu report's stall samples / executed instructions to CUDA source lines.
usage: tools/ncu_hot.py report.ncu-rep <kernel-substr> <cubin-name e.g. stream> [N]
Works offline: SASS offsets from `ncu --page source` are joined with `nvdisasm -g` line info of the
cubin extracted from tantivy_aggregations_b200/libtagg.so (must be the same build)."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kern, cub = sys.argv[1], sys.argv[2], sys.argv[3]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "tantivy_aggregations_b200", "libtagg.so")], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f"{cub}.sm_100a.cubin")], capture_output=True, text=True).stdout
def parse_lines(kern):
  line_of, cur, infn = {}, None, False
  for l in dis.splitlines():
    if l.startswith("//---------------------"): infn = kern in l
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m: line_of[int(m.group(1), 16)] = cur
  return line_of
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# exact function: "void k_stream<(int)1, (int)1, (int)0, (bool)1, (bool)1>(SParams)" -> _Z8k_streamILi1ELi1ELi0ELb1ELb1EEv7SParams
demangled = rows[0][1]
mt = re.match(r"void (\w+)<(?:Shp<)?(.*?)>+\(", demangled)
if mt:
    args = "".join(("Li" if "int" in a else "Lb") + a.split(")")[1].strip().replace("-", "n") + "E" for a in mt.group(2).split(","))
    kern = f"_Z{len(mt.group(1))}{mt.group(1)}I" + (f"3ShpI{args}E" if "Shp<" in demangled else args) + "E"
    print("function", kern)
line_of = parse_lines(kern)
hi = next(i for i, r in enumerate(rows) if len(r) > 1 and r[1] == "Source")
ix = {h: i for i, h in enumerate(rows[hi])}
S, I = ix["# Samples"], ix["Instructions Executed"]
stall_cols = [(h[6:], i) for h, i in ix.items() if h.startswith("stall_") and "Not Issued" not in h]
base = None
per = collections.defaultdict(lambda: [0, 0, collections.Counter()])
ts = ti = 0
for r in rows[hi + 1:]:
    if len(r) <= max(S, I) or not r[0].startswith("0x"): continue
    a = int(r[0], 16)
    if base is None: base = a
    ln = line_of.get(a - base)
    s, i = int(r[S] or 0), int(r[I] or 0)
    ts += s; ti += i
    p = per[ln]; p[0] += s; p[1] += i
    for h, c in stall_cols: p[2][h] += int(r[c] or 0)
src = {}
print(f"total samples {ts}  total warp-instructions {ti}")
for ln, (s, i, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:N]:
    text = ""
    if ln:
        f = os.path.join(root, "tantivy_aggregations_b200", "csrc", ln[0])
        if f not in src and os.path.exists(f): src[f] = open(f).read().splitlines()
        if f in src and ln[1] <= len(src[f]): text = src[f][ln[1] - 1].strip()[:90]
    top = ", ".join(f"{k}:{v}" for k, v in st.most_common(2))
    print(f"{100*s/max(ts,1):5.1f}%s {100*i/max(ti,1):5.1f}%i {str(ln[0])+':'+str(ln[1]) if ln else '?':>16} {text:90s} [{top}]")
if os.environ.get("HOT_BY_LINE"):
    print("\n# every source line with >= 0.1 % of the executed warp instructions, in source order (instr %, samples %)")
    for ln, (s, i, st) in sorted(((k, v) for k, v in per.items() if k), key=lambda kv: kv[0]):
        if 100 * i / max(ti, 1) < 0.1: continue
        text = ""
        f = os.path.join(root, "tantivy_aggregations_b200", "csrc", ln[0])
        if f not in src and os.path.exists(f): src[f] = open(f).read().splitlines()
        if f in src and ln[1] <= len(src[f]): text = src[f][ln[1] - 1].strip()[:110]
        print(f"{100*i/max(ti,1):5.2f}%i {100*s/max(ts,1):5.2f}%s {ln[0]}:{ln[1]:<5d} {text}")
