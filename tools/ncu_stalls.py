"""Warp-stall picture of ONE kernel from an `ncu --set full --import-source on` report (source page, SASS):
usage: python tools/ncu_stalls.py X.ncu-rep [top_n]   -> stall reasons over the kernel, the instructions with the most samples,
every global / local / shared memory instruction with its execution count, and the raw-page metrics of the L1 data pipe,
the SM<->L2 ports and DRAM."""
import csv, io, subprocess, sys

rep, top_n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
print(rows[0][1][:110])
H, data = rows[1], rows[2:]
col = {h: i for i, h in enumerate(H)}
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]]) for r in data)
agg = sorted(((sum(int(r[col[s]]) for r in data), s) for s in stalls), reverse=True)
print("samples", tot, "| " + ", ".join(f"{s[6:]} {100 * n / max(tot, 1):.0f}%" for n, s in agg[:8]))
for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]]))[:top_n]):
    r = data[i]
    st = sorted(((int(r[col[s]]), s[6:]) for s in stalls if int(r[col[s]]) > 0), reverse=True)[:2]
    print(f"  {i:5d} {r[col['Source']].strip()[:64]:64s} samples {r[col['# Samples']]:>5s} exec {r[col['Instructions Executed']]:>7s} {st}")
print("memory instructions:")
for i, r in enumerate(data):
    s = r[col["Source"]].strip().split()
    op = next((t for t in s[:2] if t[:3] in ("LDG", "STG", "LDL", "STL", "LDS", "STS", "UBL", "ATO", "RED")), None)
    if op and int(r[col["Instructions Executed"]]) > 1000:
        print(f"  {i:5d} {' '.join(s)[:60]:60s} exec {r[col['Instructions Executed']]:>7s}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
want = ("gpu__time_duration.sum", "sm__cycles_elapsed.max", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed_op_local_st.sum",
        "sm__sass_inst_executed_op_local_st.sum", "sm__sass_inst_executed_op_local_ld.sum")
for h, u, d in zip(rr[0], rr[1], rr[2]):
    if h in want:
        print(f"  {h} = {d} {u}")
