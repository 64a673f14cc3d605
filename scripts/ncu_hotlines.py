"""Join an ncu SASS source page with nvdisasm line info: executed instructions per source line.

    python scripts/ncu_hotlines.py rep.ncu-rep pyrayt_b200/libpyrayt_b200.so '_ZN3prt12trace_kernelILb1EEEvNS_9TraceArgsE' [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, so, func = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    ia, ie, it, iss = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    inst = []
    for r in rows[2:]:
        try:
            inst.append((int(r[ia], 16), int(r[ie]), int(r[it]), int(r[iss]), r[hdr.index("Source")]))
        except (ValueError, IndexError):
            pass
    base = inst[0][0]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if f.startswith(os.environ.get("HOT_CUBIN", "prt_kernels") + ".") and f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout
    lines = dis.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text." + func + ":"))
    cur, off2line = ("?", 0, ""), {}
    inl = ""
    for l in lines[start + 1:]:
        if l.startswith("//-----") and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3))
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m:
            off2line[int(m.group(1), 16)] = cur
    agg, tot, tots = {}, 0, 0
    for a, e, t, s, _ in inst:
        f, ln, extra = off2line.get(a - base, ("?", 0, ""))
        k = (f, ln)
        v = agg.setdefault(k, [0, 0, 0])
        v[0] += e
        v[1] += t
        v[2] += s
        tot += e
        tots += s
    src_cache = {}

    def src(f, ln):
        for d in ("pyrayt_b200/csrc", "include"):
            p = os.path.join(d, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                if 0 < ln <= len(src_cache[p]):
                    return src_cache[p][ln - 1].strip()
        return ""

    print(f"total warp-instructions {tot}, stall samples {tots}")
    if os.environ.get("HOT_RANGES"):
        # HOT_RANGES="file:lo-hi=name,..." : executed instructions / samples per named line range
        groups = {}
        spec = []
        for part in os.environ["HOT_RANGES"].split(","):
            loc, name = part.split("=")
            f, rng = loc.split(":")
            lo, hi = rng.split("-")
            spec.append((f, int(lo), int(hi), name))
        for (f, ln), (e, t, sm) in agg.items():
            name = next((n for (ff, lo, hi, n) in spec if ff == f and lo <= ln <= hi), f"other:{f}")
            g = groups.setdefault(name, [0, 0])
            g[0] += e
            g[1] += sm
        for name, (e, sm) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
            print(f"{100 * e / tot:5.1f}% inst {100 * sm / max(tots, 1):5.1f}% smp  {name}")
        return
    for (f, ln), (e, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * e / tot:5.1f}% inst {100 * s / max(tots, 1):5.1f}% smp  {f}:{ln:<4d} {src(f, ln)[:100]}")


if __name__ == "__main__":
    main()
