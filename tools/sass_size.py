"""Static code size of one kernel by source function (nvdisasm line info): tools/sass_size.py <object.o> <kernel substring>
Code size matters for kernels whose warps sit in different phases: the instruction cache (~32-40 KB) holds the union."""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj, kern = sys.argv[1], sys.argv[2]
src = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(obj), os.path.basename(obj).replace(".o", ".cu"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
# source functions: (start line, name)
funcs = []
for i, l in enumerate(open(src), 1):
    m = re.match(r"^(?:__device__|__global__|template|static|extern).*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(", l)
    if m and not l.startswith("template"):
        funcs.append((i, m.group(1)))
    mm = re.match(r"^\s+// ([A-G])\. ", l)
    if mm:
        funcs.append((i, funcs[-1][1].split(":")[0] + ":" + mm.group(1)))
funcs.sort()
def where(f, ln):
    if os.path.basename(f) != os.path.basename(src):
        return "hdr:" + os.path.basename(f)
    name = "?"
    for a, n in funcs:
        if a <= ln:
            name = n
        else:
            break
    return name
cnt = collections.Counter()
fn, line = None, ("?", 0)
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        line = (m.group(1), int(m.group(2)))
        continue
    if fn and kern in fn and re.match(r"\s+/\*[0-9a-f]+\*/", l):
        cnt[where(*line)] += 1
tot = sum(cnt.values())
print(f"{kern}: {tot} instructions, {tot * 16} bytes")
for k, v in cnt.most_common():
    print(f"{v * 16:7d} B  {k}")
