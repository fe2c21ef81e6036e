import csv, collections, sys
for f in sys.argv[1:]:
    rows=[r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
    seq=[(r[ki].split("(")[0].replace("ais::","").replace("<unnamed>::","").replace("void ",""), float(r[vi].replace(",",""))) for r in rows[1:]]
    # split into steps at init_keys preceded by non-init... find indices of 'bm25_kernel'
    starts=[max(i-1,0) for i,(n,_) in enumerate(seq) if n.startswith("bm25_slices")]
    last=seq[starts[-1]:]
    tot=sum(v for _,v in last)
    print(f, "launches in last step", len(last), "sum us %.1f"%(tot/1e3))
    agg=collections.OrderedDict()
    for n,v in last:
        agg.setdefault(n,[0,0.0]); agg[n][0]+=1; agg[n][1]+=v
    for n,(c,v) in agg.items(): print("   %-28s x%-2d %9.1f us %5.1f%%"%(n[:28],c,v/1e3,100*v/tot))
