import csv, collections, sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
agg=collections.defaultdict(lambda:[0,0.0])
for row in csv.DictReader(lines):
    try: v=float(row['Metric Value'].replace(',',''))
    except Exception: continue
    k=row['Kernel Name'][:100]; agg[k][0]+=1; agg[k][1]+=v
tot=sum(v[1] for v in agg.values())
print('total ms %.3f launches %d'%(tot/1e6, sum(v[0] for v in agg.values())))
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print('%5.2f%% %5d %8.1f us  %s'%(100*v[1]/tot, v[0], v[1]/v[0]/1e3, k))
