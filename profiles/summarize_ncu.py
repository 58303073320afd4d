"""Turn an `ncu --set full` capture (.ncu-rep, brought back from the GPU box in gpurun_out/) into the text summary committed under
profiles/:   python profiles/summarize_ncu.py gpurun_out/<name>.ncu-rep "<title>" > profiles/<name>.txt
Per profiled launch: duration, DRAM bytes, L2 hit rate, tensor-pipe / issue / warp activity, registers, occupancy limiters; then the
warp-state samples by stall reason summed over the launches (source page)."""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"), ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active % (of active cycles)"),
        ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions (tcgen05.mma / TMA issue)"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "registers / thread"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("launch__occupancy_limit_shared_mem", "occupancy limit (smem)"), ("launch__occupancy_limit_registers", "occupancy limit (registers)"),
        ("smsp__inst_executed.sum", "instructions executed")]
print(title)
print()
for r in body:
    print(f"launch {r[col['ID']]}: {r[col['Kernel Name']]}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}")
    for k, name in KEYS:
        if k in col:
            print(f"    {name:52s} {r[col[k]]} {units[col[k]]}")
stall = Counter()
for r in body:
    for h, i in col.items():
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            try:
                stall[h[len("smsp__pcsamp_warps_issue_stalled_"):]] += float(r[i])
            except ValueError:
                pass
tot = sum(stall.values())
if tot:
    print("\nwarp-state samples by stall reason (all profiled launches):")
    for k, v in stall.most_common():
        print(f"    stall_{k:28s} {int(v):9d}  {100 * v / tot:5.1f}%")
