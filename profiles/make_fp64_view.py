"""Refresh entries of profiles/fp64_view.json (and the per-kernel text summaries) from an
ncu --set full report:
    python profiles/make_fp64_view.py gpurun_out/prof_r2_sim.ncu-rep r2
Every kernel of the report (first launch of each) replaces its entry; the others stay."""
import csv, io, json, os, re, subprocess, sys
rep, tag = sys.argv[1], sys.argv[2]
here = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
num = lambda r, k: float(r[col[k]].replace(",", "") or 0)
path = os.path.join(here, "fp64_view.json")
view = json.load(open(path))
seen = set()
for i, r in enumerate(rows[2:]):
    name = re.search(r"(k_\w+)", r[col["Kernel Name"]]).group(1)
    if name in seen:
        continue
    seen.add(name)
    view["kernels"][name] = {
        "gpu_time_us": num(r, "gpu__time_duration.sum"),
        "fp64_pipe_active_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "active_lanes_per_warp_instr": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warp_instructions": num(r, "smsp__inst_executed.sum"),
        "registers": num(r, "launch__registers_per_thread"),
        "dram_bytes": (num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")),
        "dram_bytes_unit": rows[1][col["dram__bytes_read.sum"]],
        "grid": num(r, "launch__grid_size"), "block": num(r, "launch__block_size"),
    }
    txt = subprocess.run([sys.executable, os.path.join(here, "ncu_extract.py"), rep, str(i)],
                         capture_output=True, text=True).stdout
    open(os.path.join(here, f"{tag}_{name}_ncu.txt"), "w").write(txt)
json.dump(view, open(path, "w"), indent=1)
print({k: view["kernels"][k]["gpu_time_us"] for k in seen})
