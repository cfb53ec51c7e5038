import csv, sys, subprocess
rep=sys.argv[1]; thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.006
raw=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
data=[]; tot=0
for r in rows:
    if len(r)>8 and r[0].isdigit():
        try: n=int(r[7]); st=int(r[4] or 0)
        except: continue
        data.append((int(r[0]), r[1], n, st)); tot+=n
print('total warp instr',tot)
stt=sum(d[3] for d in data)
for l,s,n,st in data:
    if n>tot*thr or st>stt*0.02: print('%5d %9d %5.1f%% stall %5.1f%% | %s'%(l,n,100*n/tot,100*st/max(stt,1),s.strip()[:105]))
