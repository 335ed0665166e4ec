"""profiles/traffic*.json + per-kernel text summaries from an ncu --set full report.
    python profiles/make_traffic.py gpurun_out/prof_r2.ncu-rep r2 262144 [cars] [out.json]
(cars defaults to 12 and the output to traffic.json; bench.py reads traffic.json for 12 cars and
traffic_c<cars>.json otherwise)"""
import csv, io, json, os, re, subprocess, sys
rep, tag, frames = sys.argv[1], sys.argv[2], int(sys.argv[3])
cars = int(sys.argv[4]) if len(sys.argv) > 4 else 12
out_name = sys.argv[5] if len(sys.argv) > 5 else ("traffic.json" if cars == 12 else f"traffic_c{cars}.json")
here = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def num(r, k):
    return float(r[col[k]].replace(",", "") or 0)
traffic, times, fp64, issue, lanes = {}, {}, {}, {}, {}
for i, r in enumerate(rows[2:]):
    name = re.search(r"(k_\w+|f?stats_kernel|plan_fused|plan_warp)", r[col["Kernel Name"]]).group(1)
    if name in traffic:  # a second launch of the same kernel: keep the first
        continue
    rd = num(r, "dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]]
    wr = num(r, "dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]]
    traffic[name] = rd + wr
    times[name] = num(r, "gpu__time_duration.sum")
    fp64[name] = num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    issue[name] = num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
    lanes[name] = num(r, "smsp__thread_inst_executed_per_inst_executed.ratio")
    txt = subprocess.run([sys.executable, os.path.join(here, "ncu_extract.py"), rep, str(i)],
                         capture_output=True, text=True).stdout
    open(os.path.join(here, f"{tag}_{name}_ncu.txt"), "w").write(txt)
pipe = {k: v for k, v in traffic.items() if k != "stats_kernel"}
json.dump({"source": f"profiles/{tag}_*_ncu.txt: ncu --set full --clock-control none, one launch of each "
                     f"kernel over {frames} frames x {cars} cars on a gpurun B200",
           "frames_per_launch": frames, "cars_per_frame": cars, "dram_bytes_per_launch": pipe,
           "stats_kernel_dram_bytes_per_launch": traffic.get("stats_kernel"),
           "gpu_time_us_under_ncu": times,
           "fp64_pipe_active_pct": fp64, "issue_active_pct": issue, "active_lanes_per_warp_instr": lanes},
          open(os.path.join(here, out_name), "w"), indent=1)
print(json.dumps(traffic), sum(pipe.values()) / frames, "B/frame")
