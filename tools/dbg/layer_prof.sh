# ncu --set full of single layers (tools/prof_kernel.py tags) summarised on the box: per-role view + top stalls
set -u
O=gpurun_out
for T in "$@"; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:conv_t -s 2 -c 1 -f -o $O/prof_$T python tools/prof_kernel.py $T 3 > $O/ncu_$T.log 2>&1
  python tools/ncu_stalls.py $O/prof_$T.ncu-rep 0 40 > $O/stalls_$T.txt 2>&1
  python tools/dbg/ncu_roles.py $O/prof_$T.ncu-rep 0 > $O/roles_$T.txt 2>&1
  ncu -i $O/prof_$T.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[-1]
for k in ('Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__shared_mem_per_block_dynamic','launch__registers_per_thread','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active'):
    print(k, r[h.index(k)] if k in h else None)
" > $O/raw_$T.txt
done
rm -f $O/*.ncu-rep
