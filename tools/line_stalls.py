#!/usr/bin/env python
"""Per-source-line stall samples of one kernel: joins the SASS page of an ncu report (per-instruction samples and stall
reasons) with nvdisasm's line table of the library's cubin (the CSV export of ncu's CUDA-source page carries no metrics).
usage: tools/line_stalls.py <report.ncu-rep> <kernel-regex> <mangled-name-substring> [launch-index] [top]"""
import csv, glob, io, os, re, subprocess, sys, tempfile
rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(REPO, "instant_nvr_b200", "csrc", "libnvr_b200.so")], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# walk the function's listing: remember the current "//## File "...", line N" and assign it to every instruction after it
lines_of, infn, cur = [], False, None
for ln in dis.splitlines():
    if ln.startswith(".text."):
        infn = mangled in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines_of.append(cur)
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
his = [i for i, r in enumerate(rows) if "Instructions Executed" in r]
hi = his[which]
end = his[which + 1] - 1 if which + 1 < len(his) else len(rows)
hdr = rows[hi]
ii, ss = hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hi + 1:end] if len(r) > ii and r[0].startswith("0x")]
if len(body) != len(lines_of):
    print("warning: %d SASS rows in the report vs %d in the cubin listing" % (len(body), len(lines_of)), file=sys.stderr)
agg = {}
for r, lo in zip(body, lines_of):
    a = agg.setdefault(lo, {"smp": 0.0, "ins": 0.0, "st": {}})
    a["smp"] += float(r[ss] or 0); a["ins"] += float(r[ii] or 0)
    for i in stall:
        v = float(r[i] or 0)
        if v:
            a["st"][hdr[i][6:]] = a["st"].get(hdr[i][6:], 0) + v
ts, ti = sum(a["smp"] for a in agg.values()), sum(a["ins"] for a in agg.values())
print("samples %d warp-instr %d" % (ts, ti))
src_cache = {}
def src(lo):
    if lo is None:
        return ""
    for root in (os.path.join(REPO, "instant_nvr_b200", "csrc"), "/usr/local/cuda/include", "/usr/local/cuda/include/crt"):
        p = os.path.join(root, lo[0])
        if os.path.exists(p):
            L = src_cache.setdefault(p, open(p, errors="replace").read().splitlines())
            return L[lo[1] - 1].strip()[:100] if lo[1] <= len(L) else ""
    return ""
for lo, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
    st = sorted(a["st"].items(), key=lambda kv: -kv[1])[:3]
    print("%-22s smp %5.1f%% ins %5.1f%% | %-60s | %s" % ("%s:%d" % lo if lo else "?", 100 * a["smp"] / ts, 100 * a["ins"] / ti, src(lo)[:60],
                                                      " ".join("%s %d%%" % (k, 100 * v / max(a["smp"], 1)) for k, v in st)))
