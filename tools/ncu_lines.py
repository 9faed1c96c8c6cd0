"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump: share of warp instructions, share of
stall samples, average active lanes, top stall reasons.  usage: ncu_lines.py dump.csv [kernel_index] [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 45
# split into kernels: a new kernel starts when a file path repeats
kernels, cur, seen, fname = [], [], set(), None
hdr = None
for r in rows:
    if r and r[0] == "File Path":
        if r[1] in seen:
            kernels.append(cur); cur = []; seen = set()
        seen.add(r[1]); fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and r and r[0] not in ("", "Kernel Name") and r[0].isdigit():
        cur.append((fname, r))
kernels.append(cur)
K = kernels[kidx]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
def num(x):
    try: return float(x)
    except: return 0.0
tot_i = sum(num(r[ix["Instructions Executed"]]) for _, r in K)
tot_s = sum(num(r[ix["# Samples"]]) for _, r in K)
print("kernel %d: warp instructions %.4g, samples %.0f" % (kidx, tot_i, tot_s))
agg = {}
for s in stalls:
    agg[s] = sum(num(r[ix[s]]) for _, r in K)
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot_s) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
K.sort(key=lambda fr: -num(fr[1][ix["# Samples"]]))
for f, r in K[:topn]:
    i, s = num(r[ix["Instructions Executed"]]), num(r[ix["# Samples"]])
    lanes = num(r[ix["Thread Instructions Executed"]]) / i if i else 0
    top = sorted(((num(r[ix[k]]), k[6:]) for k in stalls), reverse=True)[:2]
    print("%-12s %4s instr %5.1f%% samples %5.1f%% lanes %4.1f  %-22s | %s" % (f, r[0], 100 * i / tot_i, 100 * s / tot_s, lanes,
          " ".join("%s:%.0f%%" % (k, 100 * v / s if s else 0) for v, k in top), r[1].strip()[:90]))
