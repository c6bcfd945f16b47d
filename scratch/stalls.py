import re, sys, subprocess
obj, fun = sys.argv[1], sys.argv[2]
txt = subprocess.run(['cuobjdump','-sass','-fun',fun,obj],capture_output=True,text=True).stdout
lines = txt.split('\n'); ins=[]; i=0
while i<len(lines):
    m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/',lines[i])
    if m and i+1<len(lines):
        m2=re.match(r'\s+/\* (0x[0-9a-f]+) \*/',lines[i+1]); hi=int(m2.group(1),16) if m2 else 0
        ins.append((int(m.group(1),16),m.group(2).strip(),(hi>>41)&0xf,(hi>>52)&0x3f)); i+=2
    else: i+=1
print('total',len(ins))
# split into regions at backward branches
bras=[(x[0],int(re.search(r'0x([0-9a-f]+)',x[1]).group(1),16)) for x in ins if 'BRA' in x[1] and re.search(r'0x([0-9a-f]+)',x[1])]
loops=[(t,a) for a,t in bras if t<a]
for t,a in sorted(loops):
    body=[x for x in ins if t<=x[0]<=a]
    mufu=sum(1 for x in body if 'MUFU' in x[1])
    print('loop %05x-%05x: %4d instr, stall sum %5d, mufu %3d, waits %3d, inner branches %d'%(t,a,len(body),sum(x[2] for x in body),mufu,sum(1 for x in body if x[3]),sum(1 for x in body if 'BRA' in x[1])-1))
