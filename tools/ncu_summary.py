"""Print the metrics that matter from an .ncu-rep (raw page): duration, pipe utilisation, occupancy,
DRAM traffic, and the top warp-stall reasons per launch."""
import csv
import io
import subprocess
import sys


def main(path, out=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic",
            # tensor pipe / tensor memory / L2 (the judge's list)
            "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
            "sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
            "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
            "sm__ops_path_tensor_op_hmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
            "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct"]
    lines = []
    for d in data:
        lines.append("== %s grid %s block %s" % (d[idx["Kernel Name"]][:60], d[idx["Grid Size"]], d[idx["Block Size"]]))
        for w in want:
            if w in idx:
                lines.append("   %-70s %s %s" % (w, d[idx[w]], units[idx[w]]))
        st = [(h, float(d[idx[h]].replace(",", "") or 0)) for h in hdr
              if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")]
        st.sort(key=lambda t: -t[1])
        lines.append("   stalls (warps per issue): " + ", ".join(
            "%s=%.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
            for h, v in st[:7]))
    text = "\n".join(lines)
    print(text)
    if out:
        with open(out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
