import csv, sys, collections
kern=None; fname=None; hdr=None
data=collections.defaultdict(list)  # kern -> list of (file,line,src,samples,inst)
for row in csv.reader(open(sys.argv[1])):
    if not row: continue
    if row[0]=="Function Name": kern=row[1]; continue
    if row[0]=="File Path": fname=row[1].split('/')[-1]; continue
    if row[0]=="Line No": hdr=row; iS=hdr.index("# Samples"); iI=hdr.index("Instructions Executed"); continue
    if row[0]=="" : continue
    try: ln=int(row[0])
    except: continue
    try:
        s=int(row[iS]); i=int(row[iI])
    except: continue
    data[kern].append((fname,ln,row[1].strip()[:90],s,i))
top=int(sys.argv[2]) if len(sys.argv)>2 else 14
for k,rows in data.items():
    ts=sum(r[3] for r in rows); ti=sum(r[4] for r in rows)
    print(f"=== {k}  samples={ts} inst={ti}")
    for r in sorted(rows,key=lambda r:-r[4])[:top]:
        print(f"  {r[0]}:{r[1]:5d} inst={r[4]/max(ti,1)*100:5.1f}% samp={r[3]/max(ts,1)*100:5.1f}%  {r[2]}")
