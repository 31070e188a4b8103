"""Summarise an .ncu-rep (run where ncu is installed; no GPU needed): key roofline metrics and the
top stall sites of the source page.  Usage: python tools/ncu_summary.py <report.ncu-rep> [out.txt]"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__cluster_size', 'launch__shared_mem_per_block_dynamic',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum',
        'sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu.sum',
        'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_read.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum']


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    print("kernel:", d.get("Kernel Name", ("", ""))[1], file=out)
    for k in WANT:
        if k in d:
            print(f"  {k} [{d[k][0]}] = {d[k][1]}", file=out)
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warp") and "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    stalls = sorted(((float(d[h][1].replace(',', '')), h) for h in stall_cols if d[h][1] not in ("", "no data")), reverse=True)
    print("  warp stall reasons (cycles per issue):", file=out)
    for v, h in stalls[:8]:
        print(f"    {v:8.2f}  {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')}", file=out)
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    if len(src) > 2:
        h = src[0]
        try:
            si, li = h.index("Source"), h.index("# Samples") if "# Samples" in h else h.index("Warp Stall Sampling (All Samples)")
        except ValueError:
            print("  (source page columns:", h[:12], ")", file=out)
            return
        tot = 0
        items = []
        for r in src[1:]:
            try:
                n = int(r[li].replace(',', ''))
            except (ValueError, IndexError):
                continue
            tot += n
            items.append((n, r[si].strip()[:110]))
        items.sort(reverse=True)
        print(f"  top sampled instructions (of {tot} samples):", file=out)
        for n, text in items[:22]:
            print(f"    {100.0 * n / max(tot, 1):5.1f}%  {text}", file=out)


if __name__ == "__main__":
    main()
