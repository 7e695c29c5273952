#!/usr/bin/env python
"""Turn an .ncu-rep into a short markdown summary (key metrics + hottest stall sites).
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep "title" > profiles/x.md"""
import csv, io, subprocess, sys

rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print(f"# {title}\n")
print(f"Source: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none --import-source on`); per launch.\n")
for d in data[:1]:
    print(f"Kernel: `{d[hdr.index('Kernel Name')][:110]}`\n")
    print("| metric | value | unit |\n|---|---:|---|")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"| `{w}` | {d[i]} | {units[i]} |")
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 3:
    hdr, data = rows[1], rows[2:]
    si, sm, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall = [j for j, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[sm]) for r in data if len(r) > sm and r[sm].isdigit())
    tots = {hdr[j]: sum(int(r[j]) for r in data if len(r) > j and r[j].isdigit()) for j in stall}
    print("\nWarp-stall samples by reason (all sites): " +
          ", ".join(f"{k[6:]} {100*v/max(tot,1):.1f}%" for k, v in sorted(tots.items(), key=lambda kv: -kv[1])[:8]))
    print("\nHottest SASS sites by stall samples:\n\n| share | SASS | dominant stalls |\n|---:|---|---|")
    top = sorted([(int(r[sm]), i) for i, r in enumerate(data) if len(r) > sm and r[sm].isdigit()], reverse=True)[:10]
    for c, i in top:
        st = {hdr[j][6:]: int(data[i][j]) for j in stall if data[i][j].isdigit() and int(data[i][j]) > 0}
        st = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print(f"| {100*c/max(tot,1):.1f}% | `{data[i][si].strip()[:70]}` | {st} |")
