"""SASS of one kernel of an ncu --set full report in address order with per-instruction stall samples.
usage: ncu_sass_dump.py report.ncu-rep <kernel substring> [min_samples]"""
import csv, io, re, subprocess, sys
rep, want = sys.argv[1], sys.argv[2]
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(secs[:-1], secs[1:]):
    name = re.sub(r"\(int\)|\(bool\)", "", rows[a][1] if len(rows[a]) > 1 else "")
    if want not in name:
        continue
    h = rows[a + 1]
    body = [r for r in rows[a + 2:b] if len(r) == len(h)]
    iS, iSrc, iA = h.index("# Samples"), h.index("Source"), h.index("Address") if "Address" in h else 0
    iE = h.index("Warp Instructions Executed") if "Warp Instructions Executed" in h else None
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    print("##", name[:120])
    for n, r in enumerate(body):
        s = int(r[iS]) if r[iS].isdigit() else 0
        if s < mins:
            continue
        st = {h[i][6:]: int(r[i]) for i in stall if r[i].isdigit() and int(r[i]) > 0}
        ex = r[iE] if iE is not None else ""
        print(f"{n:5d} {s:6d} {ex:>9s}  {r[iSrc][:90]:90s} {st if s else ''}")
    break
