"""Condense an Nsight Compute report into the text summary kept under profiles/.

  python scripts/ncu_summary.py gpurun_out/foo.ncu-rep [units_per_launch] > profiles/foo.txt

Per profiled kernel: duration, DRAM bytes, pipe utilisation, issue rate, occupancy, and (from the source page,
when the report was taken with --import-source on) the SASS opcode mix and the warp-stall sample breakdown.
`units_per_launch` (optional) divides the instruction counts, e.g. (128-row tile x subquantizer) units of the
encode kernel.
"""
import collections
import csv
import io
import subprocess
import sys

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    raw = ncu_csv(rep, "raw")
    hdr, unit_row = raw[0], raw[1]
    print(f"# {rep}")
    names = []
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        names.append(d.get("Kernel Name", "?"))
        print(f"\n== kernel {d.get('ID', '?')}: {d.get('Kernel Name', '?')}")
        for k in RAW_KEYS:
            if k in d:
                print(f"  {k:72s} {d[k]:>18s} {unit_row[hdr.index(k)]}")
    src = ncu_csv(rep, "source")
    # the source page concatenates kernels: a "Kernel Name" row, a header row, then one row per SASS instruction
    i, kidx = 0, 0
    while i < len(src):
        if src[i] and src[i][0] == "Kernel Name":
            kname = src[i][1] if len(src[i]) > 1 else "?"
            h = src[i + 1]
            j = i + 2
            body = []
            while j < len(src) and not (src[j] and src[j][0] == "Kernel Name"):
                body.append(src[j])
                j += 1
            summarize_source(kname, h, body, units)
            i = j
            kidx += 1
        else:
            i += 1


def summarize_source(kname, h, body, units):
    try:
        i_s, i_e, i_n = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    except ValueError:
        return
    stall_cols = [(c, k) for k, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    ops, samples, stalls = collections.Counter(), collections.Counter(), collections.Counter()
    total = 0
    for r in body:
        if len(r) <= max(i_e, i_n):
            continue
        t = r[i_s].strip().split()
        if not t:
            continue
        o = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        o = o.split(".")[0]
        e = int(r[i_e] or 0)
        ops[o] += e
        samples[o] += int(r[i_n] or 0)
        total += e
        for c, k in stall_cols:
            if k < len(r) and r[k]:
                stalls[c] += int(r[k])
    print(f"\n-- SASS mix: {kname[:100]}")
    per = f" = {total / units:.1f} per unit" if units else ""
    print(f"  warp instructions executed: {total}{per}")
    for o, c in ops.most_common(24):
        pu = f" {c / units:9.1f}/unit" if units else ""
        print(f"  {o:12s} {c:14d}{pu}   {100.0 * c / max(total, 1):5.1f}%   stall samples {samples[o]}")
    ts = sum(stalls.values())
    if ts:
        print("  warp-stall samples: " + ", ".join(f"{c[6:]} {100.0 * v / ts:.1f}%" for c, v in stalls.most_common(8)))


if __name__ == "__main__":
    main()
