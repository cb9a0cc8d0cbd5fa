"""Join an ncu source-page CSV (`ncu -i X.ncu-rep --page source --csv`) with nvdisasm line info and print the
instructions executed / stall samples per SOURCE line of one kernel.
usage: ncu_by_line.py <source.csv> <object.o> <kernel-substring> [top=40]"""
import csv, os, re, subprocess, sys, tempfile, collections

src_csv, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instruction index -> source line, for the wanted kernel
lines_of = []
inside = False
cur = None
for ln in dis:
    if ln.startswith("//--------------------- .text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.match(r"\s*//## File \"([^\"]+)\", line (\d+)(.*)", ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines_of.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for i, r in enumerate(body):
    where = lines_of[i] if i < len(lines_of) else None
    inst = int(float(r[ix["Instructions Executed"]] or 0))
    samp = int(float(r[ix["# Samples"]] or 0))
    thr = int(float(r[ix["Thread Instructions Executed"]] or 0))
    agg[where][0] += inst
    agg[where][1] += samp
    agg[where][2] += thr
    tot[0] += inst
    tot[1] += samp
print("SASS instructions in report %d, in disassembly %d" % (len(body), len(lines_of)))
print("total warp instructions %d, samples %d" % (tot[0], tot[1]))
srcs = {}
for (where, (inst, samp, thr)) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if where:
        f = where[0]
        if f not in srcs:
            for root in (os.path.dirname(os.path.abspath(obj)) + "/../csrc", "."):
                pth = os.path.join(root, f)
                if os.path.exists(pth):
                    srcs[f] = open(pth).read().splitlines()
                    break
            else:
                srcs[f] = []
        if where[1] - 1 < len(srcs[f]):
            text = srcs[f][where[1] - 1].strip()[:90]
    print("%5.1f%% inst %5.1f%% samp  lanes %4.1f  %s:%s  %s" % (100.0 * inst / max(tot[0], 1), 100.0 * samp / max(tot[1], 1),
                                                               thr / max(inst, 1), where[0] if where else "?", where[1] if where else "?", text))
