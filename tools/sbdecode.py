import re,sys,subprocess
obj,kern=sys.argv[1],sys.argv[2]
lo=int(sys.argv[3],16) if len(sys.argv)>3 else 0
hi=int(sys.argv[4],16) if len(sys.argv)>4 else 1<<30
filt=sys.argv[5] if len(sys.argv)>5 else None
raw=subprocess.run(['cuobjdump','-sass',obj],capture_output=True,text=True).stdout.split('\n')
p=False;out=[]
i=0
while i<len(raw):
    l=raw[i]
    if 'Function :' in l:
        p = kern in l
    if p:
        m=re.match(r'\s+/\*([0-9a-f]{4})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/',l)
        if m and i+1<len(raw):
            m2=re.match(r'\s+/\* (0x[0-9a-f]+) \*/',raw[i+1])
            if m2:
                w1=int(m2.group(1),16)
                out.append((int(m.group(1),16),m.group(2).strip(),(w1>>41)&0xf,(w1>>46)&7,(w1>>49)&7,(w1>>52)&0x3f))
                i+=2;continue
    i+=1
print(len(out),'instructions')
for a,ins,stall,wr,rd,wait in out:
    if lo<=a<=hi and (filt is None or re.search(filt,ins) or wait):
        wm=''.join(str(b) for b in range(6) if wait>>b&1)
        print(f"{a:04x} st{stall:2d} W{wr if wr!=7 else '-'} R{rd if rd!=7 else '-'} wait[{wm:6s}] {ins[:100]}")
