"""Top stall sites (SASS, needs -lineinfo and --import-source on) of the kernels whose demangled name contains one of
the given substrings.   usage: ncu_stalls.py report.ncu-rep <substring> [<substring> ...]"""
import csv, io, re, subprocess, sys
rep, wanted = sys.argv[1], sys.argv[2:]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
seen = set()
for a, b in zip(secs[:-1], secs[1:]):
    name = re.sub(r"\(int\)|\(bool\)", "", rows[a][1] if len(rows[a]) > 1 else "")
    hit = [w for w in wanted if w in name]
    if not hit or hit[0] in seen:
        continue
    seen.add(hit[0])
    h = rows[a + 1]
    body = [r for r in rows[a + 2:b] if len(r) == len(h)]
    iS, iSrc = h.index("# Samples"), h.index("Source")
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[iS]) for r in body if r[iS].isdigit())
    agg = {}
    for r in body:
        for i in stall:
            if r[i].isdigit():
                agg[h[i][6:]] = agg.get(h[i][6:], 0) + int(r[i])
    print(f"## {name[:110]}\n   samples {tot}; stalls {dict(sorted(((k, v) for k, v in agg.items() if v), key=lambda kv: -kv[1]))}")
    for r in sorted([r for r in body if r[iS].isdigit()], key=lambda r: -int(r[iS]))[:14]:
        st = {h[i][6:]: int(r[i]) for i in stall if r[i].isdigit() and int(r[i]) > 0}
        print(f"   {r[iS]:>6s}  {r[iSrc][:72]:72s} {st}")
