"""Top CUDA source lines of an ncu report by warp-stall samples.
   python tools/ncu_top_lines.py report.ncu-rep object.o kernel_substring [N]
Maps the SASS addresses of the report's source page to source lines with nvdisasm --print-line-info."""
import csv, os, re, subprocess, sys, tempfile
rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
off2line = {}
infn = False
line = None
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = kname in m.group(1); line = None; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(\S.*?);", l)
    if m and line:
        off2line[int(m.group(1), 16)] = line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kname, "-c", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = {n: i for i, n in enumerate(rows[hi])}
data = rows[hi + 1:]
for i_, r_ in enumerate(data):          # several launches in the report: keep the first
    if r_ and r_[0] == "Kernel Name":
        data = data[:i_]
        break
data = [r_ for r_ in data if r_ and r_[0].startswith("0x")]
def num(x):
    try: return float(x)
    except Exception: return 0.0
base = min(int(r[h["Address"]], 16) for r in data)
stall_names = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
agg = {}
tot = 0.0
for r in data:
    s = num(r[h["# Samples"]]); tot += s
    key = off2line.get(int(r[h["Address"]], 16) - base, ("?", 0))
    a = agg.setdefault(key, {"s": 0.0, "inst": 0.0})
    a["s"] += s; a["inst"] += num(r[h["Instructions Executed"]])
    for n in stall_names: a[n] = a.get(n, 0.0) + num(r[h[n]])
src = {}
print("total samples %.0f, mapped instructions %d" % (tot, len(off2line)))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["s"])[:N]:
    f, ln = key
    if f not in src:
        try: src[f] = open(os.path.join(os.path.dirname(os.path.abspath(obj)), f)).read().splitlines()
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:90] if 0 < ln <= len(src[f]) else ""
    st = sorted(((a.get(n, 0), n[6:]) for n in stall_names), reverse=True)[:3]
    print(f"{a['s'] / tot * 100:5.1f}% {f}:{ln:<5d} inst {a['inst']:>9.0f}  {text:90s} " + " ".join(f"{n}:{v / max(a['s'], 1) * 100:.0f}%" for v, n in st))
