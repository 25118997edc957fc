"""Summarise an .ncu-rep (ncu --set full) per kernel launch: duration, tensor-pipe and issue utilisation, DRAM / L2 / shared-memory
traffic, occupancy and the dominant warp-stall reasons.   python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("duration_us", "gpu__time_duration.sum", "time"),
    ("sm_clock", "sm__cycles_elapsed.avg.per_second", 1),
    ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 1),
    ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", 1),
    ("sm_throughput_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("dram_read_MB", "dram__bytes_read.sum", None),
    ("dram_write_MB", "dram__bytes_write.sum", None),
    ("dram_pct_of_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("l2_throughput_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", 1),
    ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1),
    ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    ("inst_executed", "smsp__inst_executed.sum", 1),
    ("registers_per_thread", "launch__registers_per_thread", 1),
    ("grid", "launch__grid_size", 1),
    ("block", "launch__block_size", 1),
    ("local_spill_bytes_st", "smsp__inst_executed_op_local_st.sum", 1),
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h, units = rows[hdr], rows[hdr + 1]
    col = {c: i for i, c in enumerate(h)}
    print(f"# {path}")
    for r in rows[hdr + 2:]:
        if len(r) < len(h):
            continue
        print(f"\n== {r[col['Kernel Name']][:110]}")
        for name, key, scale in KEYS:
            if key in col and r[col[key]] not in ("", "n/a"):
                v = float(r[col[key]].replace(",", ""))
                u = units[col[key]]
                if scale == "time":
                    mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                    print(f"  {name:24s} {v * mult:12.2f}")
                elif scale is None:   # bytes with a unit column
                    mult = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    print(f"  {name:24s} {v * mult:12.2f}")
                else:
                    print(f"  {name:24s} {v * scale:12.2f}" + (f"   [{u}]" if scale == 1 and u else ""))
        stalls = []
        for c, i in col.items():
            if c.startswith(STALL) and c.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a"):
                stalls.append((float(r[i]), c[len(STALL):-len("_per_issue_active.ratio")]))
        stalls.sort(reverse=True)
        print("  stalls (warps per issue-active cycle): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
