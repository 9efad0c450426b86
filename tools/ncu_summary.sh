#!/bin/bash
# key metrics of a .ncu-rep capture: tools/ncu_summary.sh file.ncu-rep
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr,units=rows[0],rows[1]
keys=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","launch__grid_size","launch__block_size","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__warps_active.avg.pct_of_peak_sustained_active","sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct","smsp__inst_executed.sum","sm__cycles_elapsed.max"]
for r in rows[2:]:
    for i,h in enumerate(hdr):
        if h in keys or "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
            try:
                v=float(r[i].replace(",",""))
                if "issue_stalled" in h and v<0.05: continue
            except: pass
            print(f"{h},{r[i]},{units[i]}")
    print("----")
'
