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

# source line -> region
src = open(src_path).read().splitlines()
base = os.path.basename(src_path)
region_of = {}
name = "?"
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:static\s+)?(?:__device__|__global__|template|struct|int |void )", l)
    if m and "(" in l and not l.startswith("template"):
        mm = re.search(r"([A-Za-z_0-9]+)\s*\(", l[l.find("__") if "__launch_bounds__" not in l else 0:])
        ids = re.findall(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(", l)
        ids = [x for x in ids if x not in ("__launch_bounds__", "__device__", "__global__")]
        if ids:
            name = ids[0]
    elif l.startswith("struct ") and "{" in l:
        name = l.split()[1]
    mk = re.match(r"^\s+// ---- (.*?) -*$", l)
    if mk and name.startswith("k_project"):
        name = "k_project: " + mk.group(1)[:40]
    region_of[i] = name

def key(cur):
    if cur is None:
        return "?"
    f, ln, rest = cur
    if f == base:
        return region_of.get(ln, "?")
    for ff, l2 in re.findall(r'inlined at "([^"]+)", line (\d+)', rest):
        if os.path.basename(ff) == base:
            return region_of.get(int(l2), "?") + " (lib)"
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
