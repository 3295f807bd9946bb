"""Static SASS accounting per source line for one kernel (needs -lineinfo):
   python profiles/mov_by_line.py <lib.so> <mangled-substring> [opcode-prefix]"""
import collections, os, re, subprocess, sys, tempfile
lib, key = sys.argv[1], sys.argv[2]
pref = sys.argv[3] if len(sys.argv) > 3 else "MOV"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if "sm_100a" in f and "mhd_kernels" in f][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
inside, cur = False, None
cnt, tot = collections.Counter(), collections.Counter()
for l in dis:
    if ".section" in l and ".text." in l:
        inside = key in l
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.search(r"/\*[0-9a-f]{4}\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)", l)
    if m and cur:
        tot[cur] += 1
        op = m.group(2)
        if pref in op:
            cnt[cur] += 1
print("matching", sum(cnt.values()), "of", sum(tot.values()), "static instructions")
for k, v in cnt.most_common(30):
    print(k, v, "of", tot[k])
