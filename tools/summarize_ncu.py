"""Summarise an ncu report (--set full) into a markdown table: one row per distinct kernel (template instance), averaged
over its captured launches.  usage: python tools/summarize_ncu.py gpurun_out/prof_pass.ncu-rep > profiles/xxx.md"""
import collections, csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
M = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "smsp__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
units = dict(zip(h, rows[1]))
idx = {m: h.index(m) for m in M if m in h}
kn = h.index("Kernel Name")


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_ms(v, unit):
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "second": 1e3}.get(unit, 1)


groups = collections.OrderedDict()
for r in rows[2:]:
    if len(r) != len(h):
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("<unnamed>::", "")
    # launches of one template instance with different shapes (stem / conv / QKV / FFN ...) differ in instruction count
    inst = num(r[idx["smsp__inst_executed.sum"]])
    key = (name, float(f"{inst:.2g}"))
    groups.setdefault(key, []).append(r)

print("| kernel | launches | avg ms | DRAM read MB | DRAM write MB | tensor pipe % | issue slots % | warp instr (M) | L2 thr % | DRAM thr % | regs | grid x block |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for (name, _), rs in groups.items():
    def avg(m, conv=None):
        vals = [num(r[idx[m]]) for r in rs]
        v = sum(vals) / len(vals)
        return conv(v, units[m]) if conv else v
    print(f"| `{name}` | {len(rs)} | {avg('gpu__time_duration.sum', to_ms):.4f} | {avg('dram__bytes_read.sum', to_bytes) / 1e6:.1f} | "
          f"{avg('dram__bytes_write.sum', to_bytes) / 1e6:.1f} | {avg('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
          f"{avg('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {avg('smsp__inst_executed.sum') / 1e6:.1f} | "
          f"{avg('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {avg('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
          f"{rs[0][idx['launch__registers_per_thread']]} | {rs[0][idx['launch__grid_size']]} x {rs[0][idx['launch__block_size']]} |")
