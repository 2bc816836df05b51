"""Decodes the control word of every SASS instruction in a `cuobjdump -sass [-fun NAME] file.o` listing:
stall count, write / read dependency barrier and wait mask — shows which instruction waits for which
load.  Usage: python scripts/sass_waits.py listing.txt [first_addr_hex] [last_addr_hex]"""
import re,sys
lines=open(sys.argv[1]).read().splitlines()
out=[]
i=0
while i<len(lines):
    m=re.match(r"\s*/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* 0x([0-9a-f]{16}) \*/",lines[i])
    if m and i+1<len(lines):
        m2=re.match(r"\s*/\* 0x([0-9a-f]{16}) \*/",lines[i+1])
        if m2:
            hi=int(m2.group(1),16)
            stall=(hi>>41)&0xf; yld=(hi>>45)&1; wb=(hi>>46)&7; rb=(hi>>49)&7; wait=(hi>>52)&0x3f
            out.append((int(m.group(1),16),m.group(2).strip(),stall,wb,rb,wait))
            i+=2; continue
    i+=1
lo=int(sys.argv[2],16) if len(sys.argv)>2 else 0
hi_=int(sys.argv[3],16) if len(sys.argv)>3 else 1<<30
for a,t,stall,wb,rb,wait in out:
    if lo<=a<=hi_:
        w=''.join(str(b) for b in range(6) if wait>>b&1)
        print(f"{a:05x} st{stall:2d} W{'-' if wb==7 else wb} R{'-' if rb==7 else rb} wait[{w:6s}] {t[:95]}")
