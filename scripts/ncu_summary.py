import csv,sys
src, dst = sys.argv[1], sys.argv[2]
rows=list(csv.reader(open(src)))
hdr=rows[0]
want=['Kernel Name','Grid Size','Block Size','gpu__time_duration.sum','l1tex__m_xbar2l1tex_read_bytes.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__issue_active.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor']
want=[w for w in want if w in hdr]
idx=[hdr.index(w) for w in want]
w=csv.writer(open(dst,'w'))
w.writerow(want); w.writerow([rows[1][i] for i in idx])
for r in rows[2:]: w.writerow([r[i] for i in idx])
print(len(rows)-2,'kernels ->',dst)
