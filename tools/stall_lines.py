import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 45
hi=[i for i,r in enumerate(rows) if r and r[0]=="Address"][0]
hdr=rows[hi]
ia,isrc,ie,ist=(hdr.index(k) for k in ("Address","Source","Instructions Executed","Warp Stall Sampling (All Samples)"))
out=[]
for n,r in enumerate(rows[hi+1:]):
    try: out.append((n,r[isrc],int(r[ie]),int(r[ist])))
    except: pass
tot=sum(o[2] for o in out); ts=sum(o[3] for o in out)
print("instr",tot,"samples",ts,"static",len(out))
for n,s,e,st in sorted(out,key=lambda o:-o[3])[:top]:
    print("%4d %8d %6d %5.1f%%  %s"%(n,e,st,100*st/ts,s[:90]))
