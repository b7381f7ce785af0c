"""Attribute the per-SASS-instruction metrics of an `ncu --page source --csv --print-source sass` export to
source functions (and to the marked sections of the k_project body), using nvdisasm line info of the SAME build.
usage: ncu_regions.py <lib.so> <sass.csv> <source.cu> [kernel-substring] [queries]"""
import collections, csv, os, re, subprocess, sys, tempfile

so, sass_csv, src_path = sys.argv[1:4]
kern = sys.argv[4] if len(sys.argv) > 4 else "k_projectILb0ELi3ELb0"
nq = float(sys.argv[5]) if len(sys.argv) > 5 else 132204.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = None
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin") and "sm_100" in f:
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kern in out:
            dis = out.splitlines()
            break
assert dis, "kernel not found"
start = next(i for i, l in enumerate(dis) if ".section" in l and kern in l and ".text" in l and ".rel" not in l and ".nv" not in l)
insts, cur = [], None
for l in dis[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((m.group(2), cur))
rows = list(csv.reader(open(sass_csv)))
hdr, data = rows[1], rows[2:]
assert len(data) == len(insts), (len(data), len(insts))

# source line -> region, for the given source and every other file of its directory (headers with inlined code)
def region_map(path):
    out, name = {}, "?"
    for i, l in enumerate(open(path).read().splitlines(), 1):
        m = re.match(r"^(?:static\s+)?(?:__device__|__global__|template|struct|int |void )", l)
        if m and "(" in l and not l.startswith("template"):
            ids = re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(", l)
            ids = [x for x in ids if x not in ("__launch_bounds__", "__device__", "__global__")]
            if ids:
                name = ids[0]
        elif l.startswith("struct ") and "{" in l:
            name = l.split()[1]
        mk = re.match(r"^\s+// ---- (.*?) -*$", l)
        if mk and (name.startswith("k_project") or name.startswith("query_tail")):
            name = name.split(":")[0] + ": " + mk.group(1)[:40]
        out[i] = name
    return out

src_dir = os.path.dirname(os.path.abspath(src_path))
maps = {f: region_map(os.path.join(src_dir, f)) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))}

def key(cur):
    if cur is None:
        return "?"
    f, ln, rest = cur
    if f in maps:
        return maps[f].get(ln, "?")
    for ff, l2 in re.findall(r'inlined at "([^"]+)", line (\d+)', rest):
        if os.path.basename(ff) in maps:
            return maps[os.path.basename(ff)].get(int(l2), "?") + " (lib)"
    return f

ie, smp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
reg, regs, regstall = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
stat, hot = collections.Counter(), collections.Counter()
for (op, cur), d in zip(insts, data):
    k = key(cur)
    reg[k] += int(d[ie]); regs[k] += int(d[smp]); stat[k] += 1; hot[k] += 1 if int(d[ie]) > 0.02 * nq else 0
    for c in stall_cols:
        v = int(d[c] or 0)
        if v:
            regstall[k][hdr[c]] += v
ti, ts = sum(reg.values()), sum(regs.values())
print(f"kernel {kern}: {ti} warp instructions ({ti / nq:.0f} per query), {ts} stall samples")
tot_st = collections.Counter()
for k in regstall:
    tot_st.update(regstall[k])
print("stall mix: " + ", ".join(f"{a[6:]} {100 * b / ts:.1f}%" for a, b in tot_st.most_common(7)))
print(f"static SASS instructions {sum(stat.values())}, of which executed by > 2 % of the queries: {sum(hot.values())} ({16 * sum(hot.values()) / 1024:.0f} KB)")
print(f"{'region':44s} {'inst%':>6s} {'inst/q':>7s} {'smpl%':>6s} {'static':>6s} {'hot':>5s}  top stalls (% of all samples)")
for n, v in reg.most_common(34):
    st = ", ".join(f"{a[6:]}:{100 * b / ts:.1f}" for a, b in regstall[n].most_common(3))
    print(f"{n[:44]:44s} {100 * v / ti:6.2f} {v / nq:7.1f} {100 * regs[n] / ts:6.2f} {stat[n]:6d} {hot[n]:5d}  {st}")

# per source line detail of the regions named in PLO_REGION_LINES (comma-separated substrings)
want = [w for w in os.environ.get("PLO_REGION_LINES", "").split(",") if w]
if want:
    per_line = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for (op, cur), d in zip(insts, data):
        k = key(cur)
        if any(w in k for w in want) and cur is not None:
            per_line[(cur[0], cur[1], k)] += int(d[ie])
            ops[(cur[0], cur[1], k)][op.split()[0] if not op.startswith("@") else op.split()[1]] += int(d[ie])
    for (f, ln, k), v in per_line.most_common(45):
        text = ""
        pth = os.path.join(src_dir, f)
        if os.path.exists(pth):
            text = open(pth).read().splitlines()[ln - 1].strip()[:90]
        print(f"{v / nq:7.1f}  {f}:{ln:<5d} {text}   [{', '.join(f'{a}:{b / nq:.1f}' for a, b in ops[(f, ln, k)].most_common(4))}]")
