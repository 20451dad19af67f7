import csv, re, sys
path = sys.argv[1]
lines=[l for l in open(path) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
agg={}; tot=0
for row in rows:
    t=float(row['Metric Value'].replace(',',''))/1e3
    short=re.sub(r'\(.*','',row['Kernel Name'])[:95]
    a=agg.setdefault(short,[0,0.0]); a[0]+=1; a[1]+=t; tot+=t
mine=sum(t for k,(n,t) in agg.items() if 'pn::' in k)
print(f"launches {len(rows)} total {tot/1e3:.2f} ms; pn:: {mine/1e3:.2f} ms")
for k,(n,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 40]:
    print(f"{t:9.1f} us {n:4d} {t/n:8.1f}  {k}")
