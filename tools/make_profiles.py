"""Turn the ncu captures under gpurun_out/ into the committed evidence under profiles/:
  profiles/<round>_<name>_ncu.txt   key raw metrics + top stall sites of a `--set full` capture
  profiles/traffic.json             {bench kernel name: {"dram_bytes_per_launch", "ncu_us", ...}}
  profiles/<round>_launch_shares.txt + <round>_launches.csv   the per-launch list of the bench command
usage: python tools/make_profiles.py r1
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
CAPTURES = {   # capture file -> bench.py kernel name
    f"{rnd}_prof_b128.ncu-rep": "umma_bottleneck128",
    f"{rnd}_prof_b64.ncu-rep": "umma_bottleneck64",
}


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def num(v):
    return float(v.replace(",", ""))


traffic = {}
tp = os.path.join(P, "traffic.json")
if os.path.isfile(tp):
    traffic = json.load(open(tp))
for f, kname in CAPTURES.items():
    rep = os.path.join(G, f)
    if not os.path.isfile(rep):
        continue
    hdr, units, rows = raw_rows(rep)
    r = rows[0]
    def col(name):
        i = hdr.index(name)
        v = num(r[i])
        u = units[i]
        scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u, 1.0)
        return v * scale
    rd, wr, us = col("dram__bytes_read.sum"), col("dram__bytes_write.sum"), col("gpu__time_duration.sum")
    traffic[kname] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "ncu_us": us,
                      "kernel": r[hdr.index("Kernel Name")], "capture": f"profiles/{f[:-8]}_ncu.txt"}
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "30"],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, f[:-8] + "_ncu.txt"), "w").write(
        f"# ncu --set full --clock-control none --import-source on, python tools/prof_run.py 256 2 (bs 256, B200)\n" + txt)
    print(kname, f"{(rd + wr) / 1e6:.1f} MB dram per launch, {us:.1f} us under ncu")
json.dump(traffic, open(tp, "w"), indent=1)
ll = os.path.join(G, f"{rnd}_launches.csv")
if os.path.isfile(ll):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_shares.py"), ll,
                    os.path.join(P, f"{rnd}_launch_shares.txt")], stdout=subprocess.DEVNULL)
    import shutil
    shutil.copy(ll, os.path.join(P, f"{rnd}_launches.csv"))
