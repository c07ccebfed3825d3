#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the small text files kept under profiles/.

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_autorally   -> *_kernels.csv, *_stalls.txt
"""
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
       "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sass__inst_executed_local_loads",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
STALLS = ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_selected", "stall_not_selected", "stall_math",
          "stall_mio", "stall_lg", "stall_tex", "stall_barrier", "stall_branch_resolving", "stall_dispatch",
          "stall_no_inst", "stall_membar", "stall_sleep"]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    with open(prefix + "_kernels.csv", "w") as f:
        w = csv.writer(f)
        cols = [c for c in RAW if c in hdr]
        w.writerow(["kernel"] + [f"{c} [{units[hdr.index(c)]}]" for c in cols])
        for r in rows[2:]:
            w.writerow([r[hdr.index("Kernel Name")][:70]] + [r[hdr.index(c)] for c in cols])
    src = ncu(rep, "source")
    secs = [i for i, r in enumerate(src) if r and r[0] == "Kernel Name"]
    with open(prefix + "_stalls.txt", "w") as f:
        seen = set()
        for si, i0 in enumerate(secs):
            name = src[i0][1][:90]
            if name in seen:
                continue
            seen.add(name)
            i1 = secs[si + 1] if si + 1 < len(secs) else len(src)
            h, data = src[i0 + 1], src[i0 + 2:i1]
            if "# Samples" not in h:
                continue
            isamp, isrc, iex = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
            tot = sum(int(r[isamp] or 0) for r in data)
            f.write(f"== {name}\n   sampled warps {tot}, SASS lines {len(data)}, "
                    f"warp instructions {sum(int(r[iex] or 0) for r in data)}\n")
            agg = {n: sum(int(r[h.index(n)] or 0) for r in data) for n in STALLS if n in h}
            f.write("   stall mix: " + ", ".join(f"{k[6:]} {100.0 * v / max(tot, 1):.1f}%" for k, v in
                                                  sorted(agg.items(), key=lambda x: -x[1]) if v) + "\n")
            for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:12]:
                st = {n: int(r[h.index(n)] or 0) for n in STALLS if n in h}
                f.write(f"   {r[isamp]:>6} {max(st, key=st.get)[6:]:<14} {r[isrc][:100]}\n")
    print("wrote", prefix + "_kernels.csv", prefix + "_stalls.txt")


if __name__ == "__main__":
    main()
