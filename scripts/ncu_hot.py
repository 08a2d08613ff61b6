"""Top stall sites of one kernel from `ncu -i X.ncu-rep --page source --csv` (SASS view).  usage: ncu_hot.py src.csv [n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
body = [r for r in rows[hi + 1:] if len(r) >= len(h) - 2]
def num(r, k):
    try: return float(r[ix[k]] or 0)
    except Exception: return 0.0
tot = sum(num(r, "# Samples") for r in body) or 1
print("total samples", tot, "instructions", len(body))
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
for r in sorted(body, key=lambda r: -num(r, "# Samples"))[:n]:
    top = sorted(((num(r, k), k) for k in stalls), reverse=True)[:2]
    print("%6.1f%% %-70s exec %-9s %s" % (100 * num(r, "# Samples") / tot, r[ix["Source"]][:70], r[ix["Instructions Executed"]],
                                       " ".join("%s=%d" % (k[6:], v) for v, k in top if v)))
