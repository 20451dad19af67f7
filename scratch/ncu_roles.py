"""Stall samples of a warp-specialised kernel grouped by mbarrier wait site and by code region (SASS order)."""
import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]; data = rows[2:]
ix = {k: i for i, k in enumerate(h)}
def f(r, k):
    try: return float(r[ix[k]].replace(',', ''))
    except Exception: return 0.0
acc = 0
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 60
for i, r in enumerate(data):
    src = r[ix['Source']]
    s = f(r, '# Samples'); acc += s
    key = any(t in src for t in ('TRYWAIT', 'UTMALDG', 'UTCHMMA', 'UTCBAR', 'LDTM', 'BAR.SYNC', 'UTMASTG', 'EXIT', 'ARRIVE'))
    if key or s > thr:
        print(f"{i:5d} s={s:6.0f} n={f(r,'Instructions Executed'):10.0f} cum={acc:7.0f} {src[:100]}")
print('total samples', acc)
