"""Join an ncu SASS-level source page with nvdisasm line info and aggregate per
source line / per device function.

    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep <kernel-substring> [top]
    PP_SECTION=<mangled-substring> ... when the demangled name ncu shows ("k_decide_t<0>")
    and the ELF section name (".text._ZN..10k_decide_tILb0EEE...") need different substrings

Needs ncu, cuobjdump, nvdisasm (CPU box is fine) and the libpp_b200.so the
report was captured from (built with -lineinfo).
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "carnd-path-planning-project_b200", "libpp_b200.so")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
addr2line = {}
for f in os.listdir(tmp):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside, cur, matched = False, None, False
    for ln in txt.splitlines():
        if ln.startswith(".text."):
            # the first matching section only: template instances (k_emit<PairOut>, <ArrayOut>)
            # share a name and their offsets would overwrite each other
            inside = (os.environ.get("PP_SECTION", kern) in ln) and not matched
            matched = matched or inside
            cur = None
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
        if m:
            addr2line[int(m.group(1), 16)] = (cur, m.group(2))
    if addr2line:
        break

raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
allrows = list(csv.reader(io.StringIO(raw)))
# a report may hold several kernels: keep the section whose name matches
starts = [i for i, r in enumerate(allrows) if r and r[0] == "Kernel Name"]
sel = [i for i in starts if kern in allrows[i][1]]
lo = sel[0] if sel else starts[0]
hi = min([i for i in starts if i > lo] + [len(allrows)])
rows = allrows[lo:hi]
h = rows[1]
ia, isamp, iinst, ithr = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
ino = h.index("stall_no_inst")
base = None
per_line = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0, 0]
for r in rows[2:]:
    if len(r) <= ithr or not r[ia]:
        continue
    a = int(r[ia], 16) if not r[ia].isdigit() else int(r[ia])
    if base is None:
        base = a
    key = addr2line.get(a - base, (None, "?"))[0]
    v = [int(float(r[isamp] or 0)), int(float(r[iinst] or 0)), int(float(r[ithr] or 0)), int(float(r[ino] or 0))]
    for i in range(4):
        per_line[key][i] += v[i]
        tot[i] += v[i]

# function ranges from the source files (PPD_INLINE / __device__ / __global__ definitions)
def func_ranges(path):
    out, name, start = [], None, None
    for i, ln in enumerate(open(path), 1):
        m = re.match(r"^(?:PPD_INLINE|__device__|__global__)[^;]*?\b(\w+)\s*\(", ln)
        if m and not ln.strip().endswith(";"):
            if name:
                out.append((start, i - 1, name))
            name, start = m.group(1), i
    if name:
        out.append((start, 10**9, name))
    return out
ranges = {}
for fn in ("pp_device.cuh", "pp_plan.cu"):
    ranges[fn] = func_ranges(os.path.join(ROOT, "carnd-path-planning-project_b200", "csrc", fn))
per_func = collections.defaultdict(lambda: [0, 0, 0, 0])
for key, v in per_line.items():
    fname = "(libdevice / no line)"
    if key and key[0] in ranges:
        for lo, hi, nm in ranges[key[0]]:
            if lo <= key[1] <= hi:
                fname = nm
    elif key:
        fname = key[0]
    for i in range(4):
        per_func[fname][i] += v[i]

print(f"total samples {tot[0]}  warp-instr {tot[1]}  thread-instr {tot[2]}  avg lanes {tot[2]/max(tot[1],1):.1f}  no_inst samples {tot[3]}")
print("\n-- by function: samples%  warp-instr%  avg-lanes  no_inst% of its samples")
for nm, v in sorted(per_func.items(), key=lambda t: -t[1][0]):
    print(f"{nm:34s} {100*v[0]/tot[0]:6.1f}  {100*v[1]/tot[1]:6.1f}  {v[2]/max(v[1],1):5.1f}  {100*v[3]/max(v[0],1):5.1f}")
print(f"\n-- top {top} source lines by samples")
for key, v in sorted(per_line.items(), key=lambda t: -t[1][0])[:top]:
    print(f"{str(key):34s} {100*v[0]/tot[0]:6.2f}%  inst {100*v[1]/tot[1]:6.2f}%  lanes {v[2]/max(v[1],1):5.1f}  no_inst {100*v[3]/max(v[0],1):5.1f}%")

# lane-occupancy histogram of the instruction stream
buckets = collections.OrderedDict((("1-4", 0), ("5-12", 0), ("13-24", 0), ("25-30", 0), ("31-32", 0)))
for key, v in per_line.items():
    if not v[1]:
        continue
    l = v[2] / v[1]
    b = "1-4" if l <= 4.5 else "5-12" if l <= 12.5 else "13-24" if l <= 24.5 else "25-30" if l <= 30.5 else "31-32"
    buckets[b] += v[1]
print("\n-- warp instructions by average active lanes of their source line")
for b, n in buckets.items():
    print(f"  lanes {b:6s} {100*n/tot[1]:6.1f} %")
