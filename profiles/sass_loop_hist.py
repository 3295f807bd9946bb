import re,sys,collections
def hist(fn, lo=None, hi=None):
    c=collections.Counter()
    for ln in open(fn):
        m=re.match(r'\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ln)
        if not m: continue
        a=int(m.group(1),16)
        if lo is not None and not (lo<=a<hi): continue
        op=m.group(3).split('.')[0]
        c[op]+=1
    return c
a=hist(sys.argv[1], int(sys.argv[2],16), int(sys.argv[3],16))
b=hist(sys.argv[4], int(sys.argv[5],16), int(sys.argv[6],16))
print(sum(a.values()), sum(b.values()))
for k in sorted(set(a)|set(b), key=lambda k:-(a[k]+b[k])):
    print(f"{k:12s} {a[k]:6d} {b[k]:6d}")
