"""Per-launch summary of an `ncu --set full` capture of the tcgen05 kernels: duration, tensor-pipe activity, the share of
the shared-memory datapath taken by tensor-core operand reads, L1/LSU wavefronts, DRAM bytes, L2 hit rate.
Usage: python profiles/ncu_conv_summary.py <file.ncu-rep>"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [("gpu__time_duration.sum", "duration"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active %"),
        ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor memory (TMEM) active %"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem datapath: tensor-core operand wavefronts %"),
        ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "smem/L1 datapath: LSU wavefronts %"),
        ("l1tex__data_pipe_lsu_wavefronts.sum", "LSU wavefronts"),
        ("sm__inst_executed_pipe_tc.sum", "tcgen05.mma instructions (pipe_tc)"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
        ("lts__t_bytes.sum", "L2 bytes"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "registers"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem"), ("sm__cycles_elapsed.max", "cycles")]
kn = hdr.index("Kernel Name")
for r in rows[2:]:
    print("-----", r[kn][:110])
    for key, label in want:
        if key in hdr:
            i = hdr.index(key)
            print("  %-52s %s %s" % (label, r[i], units[i]))
