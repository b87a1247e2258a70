"""Turn the ncu captures under gpurun_out/ into the committed evidence under profiles/:
  profiles/<round>_step_ncu.txt      per-kernel key metrics of one `--set full` capture of a whole pipeline step
                                     (python tools/prof_run.py 256 1) + top stall sites of the dominant kernels
  profiles/<round>_contour_ncu.txt   the same for the eight contour_noise_removal kernels (tools/bench_contour.py)
  profiles/traffic.json              {bench kernel name: {"dram_bytes_per_launch", "ncu_us", "launches", ...}}
  profiles/<round>_launch_shares.txt + <round>_launches.csv   the per-launch list of the bench command
usage: python tools/make_profiles.py r1
"""
import csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

RULES = [   # (regex on the demangled kernel name, bench.py kernel label)
    (r"k_umma_initial", "umma_initial"),
    (r"k_umma_down<\(?int\)?16>|k_umma_down<16>", "umma_pool_conv16"),
    (r"k_umma_down<\(?int\)?64>|k_umma_down<64>", "umma_pool_conv64"),
    (r"k_umma_bottleneck<64, 16, 16, 16", "umma_down64"),
    (r"k_umma_bottleneck<128, 16, 32, 64", "umma_down128"),
    (r"k_umma_bottleneck<64, 16, 16, 64", "umma_bottleneck64"),
    (r"k_umma_bottleneck<128, 32, 32, 128, 2, 2, 1", "umma_conv5x1"),
    (r"k_umma_bottleneck<128, 32, 32, 128, 2, 1, 0, 1, 1>", "umma_asym_fused"),
    (r"k_umma_bottleneck<128, 32, 32, 128", "umma_bottleneck128"),      # regular / dilated and the 1x5 halves
    (r"k_umma_up<128", "umma_up4"),
    (r"k_umma_up<64", "umma_up5"),
    (r"k_stage5_bottleneck", "stage5_bottleneck"),
    (r"k_umma_head", "umma_head_argmax_lut"),
    (r"k_occgrid", "occgrid"),
    (r"k_cn_(\w+)", None),          # contour kernels keep their own names
]
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size"]
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3}


def label(name):
    name = re.sub(r"\(int\)|\(bool\)", "", name)
    for rx, lab in RULES:
        m = re.search(rx, name)
        if m:
            return lab if lab else "cn_" + m.group(1)
    return None


def capture(rep, out_txt, header, stall_reps=()):
    if not os.path.isfile(rep):
        return {}
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    agg = {}
    for r in body:
        lab = label(r[ki])
        if lab is None:
            continue
        a = agg.setdefault(lab, {"launches": 0, "kernel": r[ki]})
        a["launches"] += 1
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                try:
                    v = float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)
                except ValueError:
                    continue
                a[w] = a.get(w, 0.0) + v
    lines = [header, f"{'kernel':24s} {'n':>3s} {'us':>8s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'dram %':>7s} {'L2 %':>6s} "
                     f"{'SM %':>6s} {'issue %':>7s} {'warps %':>7s} {'Minst':>7s} {'regs':>5s}"]
    res = {}
    for lab, a in sorted(agg.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
        n = a["launches"]
        g = lambda k: a.get(k, 0.0) / n
        lines.append(f"{lab:24s} {n:3d} {g(WANT[0]):8.1f} {g(WANT[1]) / 1e6:10.1f} {g(WANT[2]) / 1e6:10.1f} {g(WANT[3]):7.1f} "
                     f"{g(WANT[4]):6.1f} {g(WANT[5]):6.1f} {g(WANT[6]):7.1f} {g(WANT[7]):7.1f} {g(WANT[8]) / 1e6:7.2f} {g(WANT[9]):5.0f}")
        res[lab] = {"dram_bytes_per_launch": g(WANT[1]) + g(WANT[2]), "dram_read": g(WANT[1]), "dram_write": g(WANT[2]),
                    "ncu_us": g(WANT[0]), "launches": n, "kernel": a["kernel"], "capture": os.path.relpath(out_txt, ROOT)}
    txt = "\n".join(lines) + "\n"
    for srep, sub in stall_reps:
        if os.path.isfile(srep):
            txt += "\n" + subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), srep, sub],
                                         capture_output=True, text=True).stdout
    open(out_txt, "w").write(txt)
    print(txt[:3000])
    return res


traffic = {}
if os.path.isfile(os.path.join(P, "traffic.json")):            # captures that were not re-taken keep their entries
    traffic = json.load(open(os.path.join(P, "traffic.json")))
traffic.update(capture(os.path.join(G, f"{rnd}_step_light.ncu-rep"), os.path.join(P, f"{rnd}_step_ncu.txt"),
                       "# ncu --section SpeedOfLight --section MemoryWorkloadAnalysis_Tables --section LaunchStats --section Occupancy "
                       "--section SchedulerStats --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum "
                       "--clock-control none python tools/prof_run.py 256 1\n# (one pipeline step, bs 256, B200; per-launch averages; "
                       "stall sites below from --set full captures of one launch each)",
                       stall_reps=((os.path.join(G, f"{rnd}_b128_full.ncu-rep"), "k_umma_bottleneck<128, 32, 32, 128, 2, 1"),
                                   (os.path.join(G, f"{rnd}_b64_full.ncu-rep"), "k_umma_bottleneck<64, 16, 16, 64"),
                                   (os.path.join(G, f"{rnd}_init_full.ncu-rep"), "k_umma_initial_u8"),
                                   (os.path.join(G, f"{rnd}_head_full.ncu-rep"), "k_umma_head"),
                                   (os.path.join(G, f"{rnd}_up4_full.ncu-rep"), "k_umma_up<128"),
                                   (os.path.join(G, f"{rnd}_s5_full.ncu-rep"), "k_stage5"),
                                   (os.path.join(G, f"{rnd}_occ_full.ncu-rep"), "k_occgrid"))))
traffic.update(capture(os.path.join(G, f"{rnd}_contour_full.ncu-rep"), os.path.join(P, f"{rnd}_contour_ncu.txt"),
                       "# ncu --set full --clock-control none -k regex:k_cn python tools/bench_contour.py  (256 masks of 256x512, B200)"))
if traffic:
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
ll = os.path.join(G, f"{rnd}_launches.csv")
if os.path.isfile(ll):
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_shares.py"), ll,
                    os.path.join(P, f"{rnd}_launch_shares.txt")], stdout=subprocess.DEVNULL)
    import shutil
    shutil.copy(ll, os.path.join(P, f"{rnd}_launches.csv"))
