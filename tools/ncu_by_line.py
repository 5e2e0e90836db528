"""Join an `ncu --page source --csv` SASS export with nvdisasm line info and aggregate executed instructions per source line /
per enclosing function-ish region.  usage: ncu_by_line.py <ncu_sass.csv> <nvdisasm -g -c output> <mangled kernel substring> [top]"""
import collections, csv, re, sys
csv_path, sass_path, kernel = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(csv_path)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]; col = {n: i for i, n in enumerate(H)}
data = rows[hdr + 1:]
# line info per instruction, in order
lines = []
cur = ("?", 0); infn = False
for ln in open(sass_path):
    if ln.startswith("//---") and ".text." in ln:
        infn = kernel in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", ln):
        lines.append((cur, ln.split("*/", 1)[1].strip()))
assert len(lines) == len(data), (len(lines), len(data))
inst = collections.Counter(); thr = collections.Counter(); samples = collections.Counter()
tot_i = tot_t = tot_s = 0
for (loc, sass), r in zip(lines, data):
    i = int(r[col["Instructions Executed"]]); t = int(r[col["Thread Instructions Executed"]]); s = int(r[col["# Samples"]])
    inst[loc] += i; thr[loc] += t; samples[loc] += s
    tot_i += i; tot_t += t; tot_s += s
print("total warp-instructions %.3e  thread-instructions %.3e  avg active lanes %.2f  samples %d" % (tot_i, tot_t, tot_t / max(tot_i, 1), tot_s))
src_cache = {}
def src(loc):
    f, l = loc
    import glob, os
    if f not in src_cache:
        cand = glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "minimal_volumetric_path_tracer_b200", "csrc", f))
        src_cache[f] = open(cand[0]).read().split("\n") if cand else []
    L = src_cache[f]
    return L[l - 1].strip()[:90] if 0 < l <= len(L) else ""
print("%-28s %8s %8s %6s %7s  %s" % ("file:line", "inst%", "stall%", "lanes", "cum%", "source"))
cum = 0
for loc, i in inst.most_common(top):
    cum += i
    print("%-28s %7.2f%% %7.2f%% %6.1f %6.1f%%  %s" % ("%s:%d" % loc, 100 * i / tot_i, 100 * samples[loc] / max(tot_s, 1), thr[loc] / max(i, 1), 100 * cum / tot_i, src(loc)))
