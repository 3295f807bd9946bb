"""Dynamic SASS opcode histogram + top stall sites from `ncu --page source --csv` output."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Address")
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
tot = sum(int(r[iS]) for r in data)
totI = sum(int(r[iI]) for r in data)
print("total samples", tot, "warp insts", totI, "static", len(data))
h, hs = collections.Counter(), collections.Counter()
for r in data:
    s = re.sub(r"^@!?U?P\w+\s+", "", r[isrc].strip())
    op = s.split()[0].split(".")[0].rstrip(";")
    h[op] += int(r[iI])
    hs[op] += int(r[iS])
for op, c in h.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{op:10s} dyn={100*c/totI:5.1f}%  samples={100*hs[op]/max(tot,1):5.1f}%")
print()
for r in sorted(data, key=lambda r: -int(r[iS]))[:20]:
    print(r[iS].rjust(6), r[isrc].strip()[:100])
