import csv, sys, subprocess, json
want = {
 'gpu__time_duration.sum':'duration_us','sm__cycles_elapsed.avg.per_second':'sm_clock_GHz',
 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active':'tensor_pipe_active_pct',
 'dram__bytes_read.sum':'dram_read','dram__bytes_write.sum':'dram_write',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed':'dram_throughput_pct',
 'lts__throughput.avg.pct_of_peak_sustained_elapsed':'l2_throughput_pct','lts__t_sector_hit_rate.pct':'l2_hit_pct',
 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed':'l1tex_throughput_pct',
 'sm__throughput.avg.pct_of_peak_sustained_elapsed':'sm_throughput_pct',
 'sm__warps_active.avg.pct_of_peak_sustained_active':'warps_active_pct','launch__registers_per_thread':'regs',
 'launch__grid_size':'grid','launch__block_size':'block','launch__shared_mem_per_block_dynamic':'dyn_smem',
}
def summarize(path):
    out = subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    r = list(csv.reader(out.splitlines()))
    hdr, units, row = r[0], r[1], r[-1]
    d = {'kernel': row[hdr.index('Kernel Name')]}
    for h,u,v in zip(hdr,units,row):
        if h in want: d[want[h]] = f"{v} {u}".strip()
    return d
for p in sys.argv[1:]:
    print(json.dumps(summarize(p)))
