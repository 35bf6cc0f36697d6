import csv, collections, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
agg=collections.defaultdict(lambda:[0,0.0])
n=len(rows)-1
for r in rows[1+2*n//3:]:
    k=r[ki][:50]; agg[k][0]+=1; agg[k][1]+=float(r[vi].replace(",",""))
tot=sum(v[1] for v in agg.values())
print("total ms", tot/1e6)
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:5]: print("%-52s n=%3d  %9.1f us  avg %7.1f  %5.1f%%"%(k,v[0],v[1]/1e3,v[1]/1e3/v[0],100*v[1]/tot))
