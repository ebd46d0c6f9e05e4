"""Executed-instruction opcode mix of an ncu --set full --import-source on report:  python tools/ncu_opmix.py rep.ncu-rep"""
import csv, sys, subprocess, collections
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; ci={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)>5 and r[0].startswith('0x')]
mix=collections.Counter(); tot=0
for r in data:
    src=r[1].strip()
    parts=src.split()
    op=parts[1] if parts[0].startswith('@') else parts[0]
    op=op.split('.')[0]
    n=float(r[ci['Instructions Executed']] or 0)
    mix[op]+=n; tot+=n
print('total warp instr', tot)
for op,n in mix.most_common(28): print(f'{op:12s} {n/1e6:9.2f}M {100*n/tot:5.1f}%')
