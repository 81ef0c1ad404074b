#!/usr/bin/env python
"""Summarise an ncu report (read on the CPU box) into profiles/<tag>_ncu_summary.md:
    python scripts/summarize_ncu.py gpurun_out/r01a_msgpack_full.ncu-rep profiles/r01a_msgpack_ncu_summary.md
Uses `ncu -i <rep> --page raw --csv` for launch metrics and `--page source --csv` for the SASS-level
instruction mix and warp-stall samples."""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "lts__t_sectors_op_read.sum", "sm__cycles_elapsed.max"]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main(rep, out):
    lines = [f"# ncu summary of `{rep}`", ""]
    raw = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        name = r[hdr.index("Kernel Name")]
        lines += [f"## launch: `{name[:90]}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in RAW_KEYS:
            if k in hdr and r[hdr.index(k)] not in ("", "nan", "-nan"):
                lines.append(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |")
        lines.append("")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv"]))))
    h = next(i for i, r in enumerate(src) if "Address" in r)
    hd = src[h]
    ix = {k: i for i, k in enumerate(hd)}
    data = [r for r in src[h + 1:] if len(r) == len(hd)]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except ValueError:
            return 0.0
    tot_i = sum(f(r, "Instructions Executed") for r in data) or 1
    tot_s = sum(f(r, "# Samples") for r in data) or 1
    byop, sop = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[ix["Source"]].split()
        if not toks:
            continue
        op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
        byop[op] += f(r, "Instructions Executed")
        sop[op] += f(r, "# Samples")
    lines += ["## SASS instruction mix (all profiled launches)", "", "| opcode | % of warp instructions | % of stall samples |", "|---|---|---|"]
    for op, c in byop.most_common(14):
        lines.append(f"| {op} | {c / tot_i * 100:.1f} | {sop[op] / tot_s * 100:.1f} |")
    st = collections.Counter()
    for k in hd:
        if k.startswith("stall_") and "Not Issued" not in k:
            st[k] = sum(f(r, k) for r in data)
    tots = sum(st.values()) or 1
    lines += ["", "## warp stall reasons (sampled)", "", "| reason | % |", "|---|---|"]
    for k, v in st.most_common(10):
        lines.append(f"| {k} | {v / tots * 100:.1f} |")
    ti = sum(f(r, "Thread Instructions Executed") for r in data)
    lines += ["", f"average active threads per warp instruction: {ti / tot_i:.1f}", ""]
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
