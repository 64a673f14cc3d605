"""Where a kernel spills: STL/LDL instructions of one function with their source lines.
    python scripts/spills.py /tmp/kstat.o '_ZN3prt12trace_kernelILb1ELb0EEEvNS_9TraceArgsE'
"""
import re, subprocess, sys
obj, func = sys.argv[1], sys.argv[2]
dis = subprocess.run(["nvdisasm", "-g", obj], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + func + ":"))
cur, n = None, 0
for l in dis[start + 1:]:
    if l.startswith("//-----") and ".text." in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m:
        n += 1
        if "STL" in m.group(2) or "LDL" in m.group(2):
            print(m.group(1), m.group(2).strip(), cur)
print("instructions", n)
