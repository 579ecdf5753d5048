"""One eager NFE of the DiT (plus decode + render) at the benchmark shapes between
cudaProfilerStart/Stop, for an ncu launch list:

  ncu --profile-from-start off --cache-control none --clock-control none \
      --metrics gpu__time_duration.sum --csv --log-file gpurun_out/nfe.csv python tools/nfe_breakdown.py

and `python tools/nfe_breakdown.py --summarise gpurun_out/nfe.csv` to aggregate by kernel and grid.
"""
import csv
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def summarise(path, out=None):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, ig, ib, iv = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Block Size"), hdr.index("Metric Value")
    iu = hdr.index("Metric Unit")
    agg, order = {}, []
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * (1e3 if r[iu] in ("ms", "msecond") else 1.0)
        name = r[ik].split("(")[0]
        key = (name, r[ig], r[ib])
        if key not in agg:
            agg[key] = [0, 0.0]
            order.append(key)
        agg[key][0] += 1
        agg[key][1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["kernel,grid,block,launches,avg_us,total_us,share_pct"]
    for k in sorted(order, key=lambda k: -agg[k][1]):
        n, t = agg[k]
        lines.append(f'"{k[0]}","{k[1]}","{k[2]}",{n},{t / n:.2f},{t:.1f},{100 * t / tot:.2f}')
    lines.append(f'"TOTAL",,,{sum(v[0] for v in agg.values())},,{tot:.1f},100.0')
    txt = "\n".join(lines)
    if out:
        open(out, "w").write(txt + "\n")
    print(txt)


def main():
    import torch
    import bench as B
    from gvfdiffusion_b200.pipeline import GVFPipeline
    dev = torch.device("cuda", 0)
    dit, vae = B.build_models(dev, seed=0)
    pipe = GVFPipeline(dit, vae, B.reference_betas(), device=dev, resolution=B.RES)
    hin = B.host_inputs(seed=0)
    canon = {k: v.to(dev) for k, v in hin["canon"].items()}
    noise, cond = hin["noise"].to(dev), hin["cond_images"].to(dev)
    obj = pipe.prepare_object(canon)
    dit.engine().use_graphs = False
    pipe.sampler_graph = False                                      # eager launches: what a replayed object consists of
    lat = pipe.sample(obj, cond, noise, steps=2)                   # warm-up, hoisted projections
    delta = pipe.decode(lat, obj)
    pipe.render(obj, delta, hin["ext"], hin["intr"])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "nfe"):
        lat = pipe.sample(obj, cond, noise, steps=2)      # two NFEs
    if what in ("all", "tail"):
        delta = pipe.decode(lat, obj)
        pipe.render(obj, delta, hin["ext"], hin["intr"])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
    else:
        main()
