import csv,sys,subprocess,io
rep,kre=sys.argv[1],sys.argv[2]
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","-k",f"regex:{kre}"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
h=rows[1]; idx={c:i for i,c in enumerate(h)}
end=[i for i,r in enumerate(rows) if r==h]
tab=[r for r in rows[2:(end[1] if len(end)>1 else len(rows))] if len(r)==len(h)]
f=lambda r,c: float((r[idx[c]] or '0').replace(',',''))
tot_inst=sum(f(r,'Instructions Executed') for r in tab)
tot_thr=sum(f(r,'Thread Instructions Executed') for r in tab)
print('rows',len(tab),'warp inst', tot_inst, 'avg threads', tot_thr/tot_inst)
groups=[]
for r in tab:
    ie=f(r,'Instructions Executed'); th=round(f(r,'Avg. Threads Executed'),1)
    if ie==0: continue
    if groups and abs(groups[-1][0]-th)<0.05 and abs(groups[-1][3]-ie)/max(ie,1)<0.02:
        groups[-1][1]+=ie; groups[-1][2]+=1; groups[-1][5]+=f(r,'# Samples')
    else:
        groups.append([th,ie,1,ie,r[idx['Source']][:70],f(r,'# Samples')])
tots=sum(g[5] for g in groups)
thr=float(sys.argv[3]) if len(sys.argv)>3 else 0.01
for g in groups:
    if g[1]/tot_inst>thr or g[5]/tots>thr:
        print(f"lanes {g[0]:5.1f}  instr share {g[1]/tot_inst*100:5.1f}%  n {g[2]:3d}  execs {g[3]:.3g} stall share {g[5]/tots*100:5.1f}%  first: {g[4]}")
