import csv, sys, subprocess
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 30
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=[i for i,r in enumerate(rows) if 'Source' in r][0]
hdr=rows[hi]; si=hdr.index('Source'); ki=hdr.index('# Samples'); ie=hdr.index('Instructions Executed')
st=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data=[]
for n,r in enumerate(rows[hi+1:]):
    try: v=float(r[ki])
    except: continue
    data.append((v,n,r))
tot=sum(v for v,_,_ in data)
print('total samples',tot,'instructions',len(data))
for v,n,r in sorted(data,key=lambda x:-x[0])[:topn]:
    top=sorted([(float(r[i] or 0),hdr[i]) for i in st],reverse=True)[:2]
    print('%5d %6.0f %5.1f%% exec=%-9s %-70s %s' % (n, v, 100*v/tot, r[ie], r[si][:70], ' '.join('%s=%d'%(h[6:],x) for x,h in top if x>0)))
