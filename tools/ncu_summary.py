"""Summarise `ncu -i X.ncu-rep --page raw --csv` into a small JSON (per launch: duration, DRAM bytes, L2 sectors,
tensor-pipe and issue activity, registers) plus the per-step DRAM traffic table bench.py reports as `traffic`.
usage: python tools/ncu_summary.py raw.csv out_summary.json [out_traffic.json]"""
import csv
import json
import re
import sys

WANT = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "lts__t_sectors.sum": "l2_sectors",
    "sm__inst_executed.sum": "warp_instructions",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_subpipe_active_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_instructions",
    "sm__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_slots_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.max": "sm_cycles",
}
UNIT_SCALE = {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    out = []
    for r in data:
        if len(r) < len(H):
            continue
        name = re.sub(r"\(.*", "", r[H.index("Kernel Name")]).replace("void ", "").strip()
        full = r[H.index("Kernel Name")]
        m = re.search(r"<[^>]*>", full)
        e = dict(id=int(r[0]), kernel=name + (m.group(0) if m and "<" not in name else ""))
        for col, key in WANT.items():
            if col in H:
                v = r[H.index(col)].replace(",", "")
                try:
                    e[key] = float(v) * UNIT_SCALE.get(units[H.index(col)], 1.0)
                except ValueError:
                    pass
        out.append(e)
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    for e in out:
        print(f"{e['id']:4d} {e['kernel'][:58]:58s} {e.get('duration_us', 0):8.1f} us  dram {((e.get('dram_read_bytes', 0) + e.get('dram_write_bytes', 0)) / 1e6):8.1f} MB  "
              f"tensor {e.get('tensor_pipe_active_pct', e.get('tensor_subpipe_active_pct', 0)):5.1f}%  issue {e.get('issue_active_pct', e.get('issue_slots_pct', 0)):5.1f}%")
    if len(sys.argv) > 3:
        def tot(pred, per_launch):
            sel = [e for e in out if pred(e["kernel"])]
            if not sel:
                return None, 0
            b = sum(e.get("dram_read_bytes", 0) + e.get("dram_write_bytes", 0) for e in sel)
            return (b / len(sel) if per_launch else b), len(sel)

        # whole steps only: the launches from one env-step launch up to (not including) the last one captured, so that the
        # per-step sums do not depend on where in a step the capture window started
        ridx = [i for i, e in enumerate(out) if "routing_kernel" in e["kernel"]]
        if len(ridx) >= 2:
            out = out[ridx[0]:ridx[-1]]
        steps = max(1, sum(1 for e in out if "routing_kernel" in e["kernel"]))
        tc, n_tc = tot(lambda k: "linear_tc_kernel" in k or "enc_fused_kernel" in k, False)
        tr = dict(source=f"ncu --set full --clock-control none over {steps} rollout step(s) of bench.py --graph-steps 0 (cold caches, serialised), {sys.argv[1]}",
                  linear_tc_bytes_per_step=None if tc is None else tc / steps, linear_tc_launches_per_step=n_tc / steps,
                  routing_step_bytes_per_launch=tot(lambda k: "routing_kernel" in k, True)[0],
                  aggregate_pk_bytes_per_launch=tot(lambda k: "aggregate_pk" in k, True)[0],
                  readout_agents_pk_bytes_per_launch=tot(lambda k: "readout_agents" in k, True)[0],
                  replay_insert_bytes_per_step=(tot(lambda k: "replay_insert" in k, False)[0] or 0) / steps,
                  step_total_bytes=sum(e.get("dram_read_bytes", 0) + e.get("dram_write_bytes", 0) for e in out) / steps)
        json.dump(tr, open(sys.argv[3], "w"), indent=1)
        print(tr)


if __name__ == "__main__":
    main()
