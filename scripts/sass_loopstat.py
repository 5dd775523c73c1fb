import re, sys, collections
lines=[l for l in open(sys.argv[1]) if re.match(r'^\s+/\*[0-9a-f]{4,5}\*/', l)]
ins=[]
for l in lines:
    m=re.match(r'^\s+/\*([0-9a-f]+)\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1),16), m.group(2).strip()))
addr={a:i for i,(a,_) in enumerate(ins)}
loops=[]
for i,(a,t) in enumerate(ins):
    m=re.search(r'BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*0x([0-9a-f]+)', t)
    m=re.search(r'BRA.*0x([0-9a-f]+)', t)
    if m:
        tgt=int(m.group(1),16)
        if tgt<a and tgt in addr:
            body=ins[addr[tgt]:i+1]
            if sum('DFMA' in x for _,x in body) > 40: loops.append(body)
for body in loops:
    ops=collections.Counter()
    for _,t in body:
        t=re.sub(r'^@!?U?P\d+\s+','',t)
        ops[t.split()[0]]+=1
    n=len(body)
    fp=sum(v for k,v in ops.items() if k.startswith(('DFMA','DMUL','DADD')))
    print("loop: %d instrs, FP64 %d, SHFL %d"%(n,fp,sum(v for k,v in ops.items() if k.startswith('SHFL'))))
    print("   ", ops.most_common(18))
