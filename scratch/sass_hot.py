"""Top stall-sample SASS instructions of an `ncu --page source --print-source sass --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[2:]):
    try:
        s = int(r[isamp])
    except Exception:
        continue
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
    data.append((s, n, r[isrc][:70], r[iexec], top))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, n, src, ex, top in sorted(data, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{s:6d} {100*s/tot:5.1f}%  #{n:5d} exec={ex:>8} {src:70s} {top}")
