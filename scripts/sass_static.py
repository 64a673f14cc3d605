"""Static SASS instruction count per source line / function of one kernel (nvdisasm line info).

    python scripts/sass_static.py pyrayt_b200/libpyrayt_b200.so prt_kernels '_ZN3prt12trace_kernelILb1ELb0EEEvNS_9TraceArgsE' [top]
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

so, cub_prefix, func = os.path.abspath(sys.argv[1]), sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith(cub_prefix + ".") and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout
lines = dis.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + func + ":"))
cur = ("?", 0)
per_line = collections.Counter()
total = 0
for l in lines[start + 1:]:
    if l.startswith("//-----") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", l):
        per_line[cur] += 1
        total += 1
print("total static instructions", total, "=", total * 16 // 1024, "KB")
# function boundaries from the sources
funcs = {}
for fn in ("prt_device.cuh", "prt_kernels.cu", "prt_wavefront.cu", "prt_literal.cuh"):
    path = os.path.join(os.path.dirname(so), "csrc", fn)
    if not os.path.exists(path):
        continue
    marks = []
    for k, text in enumerate(open(path), 1):
        m = re.match(r"^(?:template <[^>]*>\s*)?(?:PRT_HD(?:_CALL)?|__global__|__device__|static|inline)[^;(]*?\b(\w+)\s*\(", text)
        if m and not text.startswith(" "):
            marks.append((k, m.group(1)))
    funcs[fn] = marks
per_func = collections.Counter()
for (fn, ln), n in per_line.items():
    name = "?"
    for k, nm in funcs.get(fn, []):
        if k <= ln:
            name = nm
        else:
            break
    per_func[(fn, name)] += n
for (fn, name), n in per_func.most_common(top):
    print(f"{n:6d} {100 * n / total:5.1f}%  {fn}:{name}")
