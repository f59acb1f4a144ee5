"""Summarise an .ncu-rep (read here, no GPU needed): per-launch headline metrics, stall reasons and the
hottest SASS lines.   python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex] [launch-index]"""
import csv
import io
import subprocess
import sys
from collections import Counter


def run(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def num(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


HEAD = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum"]


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else None
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(rows[2:]):
        print(f"== launch {n}: {r[idx['Kernel Name']][:90]}")
        for h in HEAD:
            if h in idx:
                print(f"   {h} = {r[idx[h]]} {units[idx[h]]}")
        st = [(h, num(r[i])) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
        st.sort(key=lambda t: -t[1])
        print("   stalls/issue:", ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for h, v in st[:8]))
    if pat is None:
        return
    args = ["-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}", "--launch-skip", str(which), "--launch-count", "1"]
    rows = list(csv.reader(io.StringIO(run(args))))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data, seen = [], set()
    for r in rows[2:]:
        if len(r) < len(hdr) - 2 or r[0] in seen:
            continue
        seen.add(r[0])
        data.append(r)
    g = lambda r, h: num(r[idx[h]])
    tot = sum(g(r, "# Samples") for r in data)
    texec = sum(g(r, "Instructions Executed") for r in data)
    print(f"\n== source page: {rows[0][1][:80]}  samples={tot:.0f} warp-instr={texec:.0f} sass-lines={len(data)}")
    c, ce = Counter(), Counter()
    for r in data:
        toks = r[idx["Source"]].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
        op = op.split(".")[0]
        c[op] += g(r, "# Samples")
        ce[op] += g(r, "Instructions Executed")
    for op, v in ce.most_common(24):
        print(f"   {op:8s} executed {ce[op]:12.0f} ({100 * ce[op] / texec:5.1f}%)   samples {c[op]:8.0f} ({100 * c[op] / tot:5.1f}%)")
    print("   hottest lines:")
    for r in sorted(data, key=lambda r: -g(r, "# Samples"))[:28]:
        print(f"   {g(r, '# Samples'):7.0f} {r[idx['Source']].strip()[:64]:64s} short={g(r, 'stall_short_sb'):.0f} mio={g(r, 'stall_mio'):.0f} "
              f"bar={g(r, 'stall_barrier'):.0f} long={g(r, 'stall_long_sb'):.0f} wait={g(r, 'stall_wait'):.0f} math={g(r, 'stall_math'):.0f}")


if __name__ == "__main__":
    main()
