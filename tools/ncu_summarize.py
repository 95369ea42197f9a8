"""Summarise an ncu report (read here, on the CPU container) into the text files kept under profiles/.

  python tools/ncu_summarize.py REPORT.ncu-rep PREFIX [--launch N]
writes  PREFIX_launches.txt      one line per profiled launch (duration, grid, instructions, issue-active, DRAM bytes)
        PREFIX_stalls.txt        warp-stall reasons per issued instruction + pipe / occupancy metrics of launch N
        PREFIX_by_function.txt   share of samples / executed instructions per device function of launch N
        PREFIX_top_lines.txt     the 40 source lines with most samples of launch N
Launch N defaults to the longest one.  Function attribution uses the source line ranges of swd_device.cuh / swd_kernels.cuh.
"""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "slidingwindowdecoder_b200", "csrc")


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def function_ranges(path):
    """[(first line, last line, name)] of the top-level functions of a .cuh file (brace counting)."""
    out, name, start, depth = [], None, 0, 0
    for i, line in enumerate(open(path), 1):
        if depth == 0:
            m = re.match(r"^(?:template.*>\s*)?(?:__device__|__global__|static|inline).*?(\w+)\s*\(", line)
            m2 = re.match(r"^(\w+)\(", line)          # kernel name on its own line after __launch_bounds__
            if m or m2:
                name, start = (m or m2).group(1), i
        depth += line.count("{") - line.count("}")
        if depth == 0 and name and "}" in line:
            out.append((start, i, name)); name = None
    return out


def main():
    rep, prefix = sys.argv[1], sys.argv[2]
    launch = int(sys.argv[sys.argv.index("--launch") + 1]) if "--launch" in sys.argv else None
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, data = rows[0], rows[2:]
    col = {k: i for i, k in enumerate(hdr)}

    def g(r, k):
        return r[col[k]] if k in col else "NA"
    dram_total = 0.0
    with open(prefix + "_launches.txt", "w") as f:
        f.write("# idx kernel duration_ms grid block regs smem_dyn_KB warp_inst issue_active_pct warps_active dram_rd_MB dram_wr_MB\n")
        for i, r in enumerate(data):
            scs = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}
            sc = scs.get(rows[1][col["dram__bytes_read.sum"]], 1.0)
            scw = scs.get(rows[1][col["dram__bytes_write.sum"]], 1.0)
            dram_total += (fnum(g(r, "dram__bytes_read.sum")) * sc + fnum(g(r, "dram__bytes_write.sum")) * scw) * 1e6
            f.write(f"{i} {g(r, 'Kernel Name')[:40]} {g(r, 'gpu__time_duration.sum')} {g(r, 'launch__grid_size')} {g(r, 'launch__block_size')} "
                    f"{g(r, 'launch__registers_per_thread')} {g(r, 'launch__shared_mem_per_block_dynamic')} {g(r, 'smsp__inst_executed.sum')} "
                    f"{g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active')} {g(r, 'sm__warps_active.avg.per_cycle_active')} "
                    f"{fnum(g(r, 'dram__bytes_read.sum')) * sc:.2f} {fnum(g(r, 'dram__bytes_write.sum')) * scw:.2f}\n")
    with open(prefix + "_launches.txt", "a") as f:
        f.write(f"# DRAM read+write over the {len(data)} launches: {dram_total:.0f} bytes = {dram_total / max(1, len(data)):.0f} per launch\n")
    if launch is None:
        launch = max(range(len(data)), key=lambda i: fnum(g(data[i], "gpu__time_duration.sum")))
    r = data[launch]
    with open(prefix + "_stalls.txt", "w") as f:
        f.write(f"# launch {launch}: {g(r, 'Kernel Name')} duration {g(r, 'gpu__time_duration.sum')} {rows[1][col['gpu__time_duration.sum']]}\n")
        f.write("# warp stall reasons, cycles per issued instruction\n")
        st = [(k.split("stalled_")[1].split("_per")[0], fnum(g(r, k))) for k in hdr if "issue_stalled" in k and "per_issue_active" in k]
        for k, v in sorted(st, key=lambda x: -x[1]):
            f.write(f"{k:24s} {v:.3f}\n")
        f.write("# other\n")
        for k in ["smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
                  "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
                  "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                  "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
                  "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                  "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
                  "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]:
            if k in col:
                f.write(f"{k} = {g(r, k)} {rows[1][col[k]]}\n")
    # ---- per source line
    txt = ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(launch), "--launch-count", "1"])
    per_line, cur_file, hdr2 = {}, None, None
    for row in csv.reader(io.StringIO(txt)):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = os.path.basename(row[1]); hdr2 = None; continue
        if row[0] == "Line No":
            hdr2 = row; continue
        if hdr2 is None or cur_file is None or not row[0].isdigit() or row[2] != "-":
            continue                                   # keep the per-line aggregate rows (Address == "-")
        c2 = {k: i for i, k in enumerate(hdr2)}
        per_line[(cur_file, int(row[0]))] = (row[1], fnum(row[c2["# Samples"]]), fnum(row[c2["Instructions Executed"]]),
                                            fnum(row[c2["Thread Instructions Executed"]]))
    tot_s = sum(v[1] for v in per_line.values()) or 1.0
    tot_i = sum(v[2] for v in per_line.values()) or 1.0
    ranges = {f: function_ranges(os.path.join(CSRC, f)) for f in ("swd_device.cuh", "swd_kernels.cuh", "swd_osd.cuh") if os.path.exists(os.path.join(CSRC, f))}
    byfn = {}
    for (f, ln), (_, s, i, t) in per_line.items():
        name = f
        for a, b, fn in ranges.get(f, []):
            if a <= ln <= b:
                name = fn; break
        e = byfn.setdefault(name, [0.0, 0.0, 0.0]); e[0] += s; e[1] += i; e[2] += t
    with open(prefix + "_by_function.txt", "w") as f:
        f.write(f"# launch {launch}; total samples {tot_s:.0f}, warp instructions {tot_i:.0f}\n")
        for name, (s, i, t) in sorted(byfn.items(), key=lambda x: -x[1][0]):
            f.write(f"{name:28s} samples {100 * s / tot_s:5.1f}%  warp-inst {100 * i / tot_i:5.1f}%  thr/inst {t / max(i, 1):4.1f}\n")
    with open(prefix + "_top_lines.txt", "w") as f:
        f.write(f"# launch {launch}; total samples {tot_s:.0f}, warp instructions {tot_i:.0f}\n")
        for (fl, ln), (src, s, i, t) in sorted(per_line.items(), key=lambda x: -x[1][1])[:40]:
            f.write(f"{fl:16s} {ln:4d}  samp {100 * s / tot_s:5.1f}%  inst {100 * i / tot_i:5.1f}%  thr/inst {t / max(i, 1):4.1f}  | {src.strip()[:110]}\n")
    print("launch", launch, "->", prefix + "_{launches,stalls,by_function,top_lines}.txt")


if __name__ == "__main__":
    main()
