"""Per-instruction warp-stall samples of an ncu --set full --import-source on report:  python tools/ncu_stalls.py rep.ncu-rep [top_n]"""
import csv, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; ci={h:i for i,h in enumerate(hdr)}
def f(x):
    try: return float(x)
    except: return 0.0
data=[r for r in rows[2:] if len(r)>5 and r[0].startswith('0x')]
tot=sum(f(r[2]) for r in data)
print('total samples',tot, 'insts', len(data))
agg={}
for r in data:
    for h in hdr:
        if h.startswith('stall_') and 'Not' not in h: agg[h]=agg.get(h,0)+f(r[ci[h]])
print(sorted(agg.items(), key=lambda kv:-kv[1])[:8])
best=sorted(enumerate(data), key=lambda x:-f(x[1][2]))[:int(sys.argv[2]) if len(sys.argv)>2 else 22]
for idx,r in best:
    st={h:r[ci[h]] for h in hdr if h.startswith('stall_') and 'Not' not in h}
    print(idx, r[2], r[1][:70].strip(), r[ci['Instructions Executed']], {k[6:]:v for k,v in st.items() if v not in ('0','')})
