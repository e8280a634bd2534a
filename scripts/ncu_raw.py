"""Print the key metrics of an `ncu --page raw --csv` export (one block per kernel launch)."""
import csv, sys
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'lts__t_sector_hit_rate.pct',
        'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum', 'l1tex__m_l1tex2xbar_write_bytes.sum',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__cycles_active.avg', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__sass_inst_executed_op_shared_ld.sum', 'sm__sass_inst_executed_op_shared_st.sum',
        'lts__t_bytes.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'dram__cycles_active.avg']
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for d in data:
        print('----', path, d[hdr.index('Kernel Name')][:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print('  %-75s %s %s' % (k, d[i], units[i]))
