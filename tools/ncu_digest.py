"""Digest of an .ncu-rep: key metrics per kernel + hottest source lines by stall samples."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_tensor_op_hmma.sum', 'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
stalls = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print("=" * 100)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:75s} {r[i]:>16s} {units[i]}")
    st = sorted(((float(r[hdr.index(h)] or 0), h) for h in stalls), reverse=True)[:7]
    for v, h in st:
        print(f"   stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:.3f}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    rs = list(csv.reader(io.StringIO(src)))
    if rs:
        h = rs[0]
        def col(name):
            for i, x in enumerate(h):
                if x.strip() == name:
                    return i
            return None
        ci, cs, cx = col("# Samples") or col("Samples"), col("Source"), col("Instructions Executed")
        print(h[:12])
        if ci is not None and cs is not None:
            body = [r for r in rs[1:] if len(r) > max(ci, cs) and r[ci].strip().isdigit()]
            tot = sum(int(r[ci]) for r in body) or 1
            for r in sorted(body, key=lambda r: -int(r[ci]))[:int(sys.argv[2])]:
                print(f"{100*int(r[ci])/tot:5.1f}%  {r[cs].strip()[:150]}")
