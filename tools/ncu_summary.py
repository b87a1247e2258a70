"""Summarise an .ncu-rep: key raw metrics per kernel + top stall instructions (needs -lineinfo, --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 22
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__inst_executed_pipe_lsu.sum"]
for r in rows[2:]:
    print("---")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w} = {r[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
for a, b in zip(secs[:1], secs[1:2]):
    h = rows[a + 1]
    body = [r for r in rows[a + 2:b] if len(r) == len(h)]
    iS, iSrc = h.index("# Samples"), h.index("Source")
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(int(r[iS]) for r in body if r[iS].isdigit())
    print("total samples", tot)
    agg = {}
    for r in body:
        for i in stall:
            if r[i].isdigit():
                agg[h[i]] = agg.get(h[i], 0) + int(r[i])
    print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    for r in sorted([r for r in body if r[iS].isdigit()], key=lambda r: -int(r[iS]))[:top_n]:
        st = {h[i][6:]: int(r[i]) for i in stall if r[i].isdigit() and int(r[i]) > 0}
        print(r[iS].rjust(6), r[iSrc][:80].ljust(80), st)
