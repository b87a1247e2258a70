"""profiles/<tag>_scale.txt from the bench lines of one scaling run: gpurun_out/<tag>_bench<N>.json, N = 1, 2, 4, 8.
usage: python tools/scale_summary.py <tag>"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
rows = []
for n in (1, 2, 4, 8):
    p = os.path.join(ROOT, "gpurun_out", f"{tag}_bench{n}.json")
    if os.path.isfile(p):
        rows.append(json.load(open(p)))
if not rows:
    sys.exit("no bench lines found")
base = rows[0]
out = [f"# python bench.py --gpus N --steps 20 --warmup 3 (torchrun for N > 1), one 8 x B200 box, bs 256 per GPU, {base['dtype']}",
       "# value = device-resident frames/s (bc_pipeline, grids peer-stored into rank 0's buffer); e2e = pinned host frames in, gathered",
       "# grids on rank 0's host, through bc_pipeline_host_submit/_wait (+ bc_gather_stream_setup at N > 1); h2d ceiling = bare",
       "# pinned-host -> device copies of the same frame batches by all ranks at once (GB/s summed over the ranks)",
       "# devices = the GPU each rank took (runtime.device_for_rank: a job smaller than the node alternates between the node's two",
       "# host domains, GPUs 0-3 and 4-7, whose H2D bandwidth is separate; the earlier placement 0,1,2,3 gave N=4 e2e 273k, ceiling 116)",
       f"{'N':>2s} {'value':>10s} {'eff':>6s} {'ms/step':>8s} {'e2e':>10s} {'eff':>6s} {'e2e ms':>8s} {'h2d need GB/s':>13s} {'h2d ceiling':>11s} {'frac':>5s} {'gather_check':>12s}  devices"]
for d in rows:
    n, e = d["n_gpus"], d["e2e"]
    out.append(f"{n:2d} {d['value']:10.0f} {d['value'] / (n * base['value'] / base['n_gpus']):6.3f} {d['ms_per_step']:8.3f} "
               f"{e['value']:10.0f} {e['value'] / (n * base['e2e']['value'] / base['n_gpus']):6.3f} {e['ms_per_step']:8.3f} "
               f"{e.get('h2d_needed_gbs', 0):13.1f} {e.get('h2d_ceiling_gbs', 0):11.1f} {e.get('e2e_frac_of_h2d_ceiling', 0):5.2f} {str(d.get('gather_check')):>12s}  {','.join(str(v) for v in d['config'].get('devices', range(n)))}")
open(os.path.join(ROOT, "profiles", f"{tag}_scale.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
