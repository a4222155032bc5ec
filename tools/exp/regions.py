#!/usr/bin/env python
"""Region table (runs of SASS instructions with similar execution counts) of the first kernel in an .ncu-rep."""
import csv, io, subprocess, math, sys
rep=sys.argv[1]; thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.01
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,text=True).stdout
rows=list(csv.reader(io.StringIO(out))); hdr=rows[1]; data=[]
for x in rows[2:]:
    if x and x[0]=="Kernel Name": break
    data.append(x)
ia,isamp,iex,ith=hdr.index("Source"),hdr.index("# Samples"),hdr.index("Instructions Executed"),hdr.index("Avg. Threads Executed")
tot_e=sum(int(x[iex]) for x in data); tot_s=sum(int(x[isamp]) for x in data)
nw=int(data[0][iex])
seg=[];start=0
def cls(e):
    e=int(e)
    return 0 if e==0 else round(math.log(e,1.6))
cur=cls(data[0][iex])
for k,x in enumerate(data):
    c=cls(x[iex])
    if c!=cur: seg.append((start,k)); cur=c; start=k
seg.append((start,len(data)))
for a,b in seg:
    e=sum(int(x[iex]) for x in data[a:b]); s=sum(int(x[isamp]) for x in data[a:b])
    if e>tot_e*thr or s>tot_s*thr:
        t=sum(float(x[ith])*int(x[iex]) for x in data[a:b])/max(1,e)
        print("%4d-%4d n=%3d exec/instr %9.0f  instr/warp %7.0f (%.1f%%) samples %.1f%% thr %.1f  first: %s"%(a,b,b-a,e/(b-a),e/nw,100*e/tot_e,100*s/tot_s,t,data[a][ia].strip()[:40]))
