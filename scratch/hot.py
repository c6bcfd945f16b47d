import csv,sys,subprocess,collections,re
rep=sys.argv[1]; rows_expected=int(sys.argv[2]) if len(sys.argv)>2 else 196608
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines())); hh=rows[1]; data=rows[2:]
ia=hh.index('Instructions Executed'); isrc=hh.index('Source')
cnt=collections.Counter(int(x[ia]) for x in data if x[ia].isdigit())
tot=sum(c*n for c,n in cnt.items())
print('total',tot,'per row',tot/rows_expected)
for c,n in sorted(cnt.items(), key=lambda kv:-kv[0]*kv[1])[:10]: print(f'count {c:9d} x {n:4d} instr = {c*n/tot*100:5.1f}%  ({c/rows_expected:.3f}/row)')
if len(sys.argv)>3:
    lo=int(sys.argv[3])
    for i,x in enumerate(data):
        if x[ia].isdigit() and int(x[ia])>=lo: print(i, x[ia], x[isrc].strip()[:95])
