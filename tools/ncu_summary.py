"""Summarise an `ncu --page raw --csv` / `--page source --csv` export: key metrics, opcode mix, stall reasons."""
import collections, csv, re, sys
raw, src = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg']
for w in want:
    for i, h in enumerate(hdr):
        if h == w or (w.startswith('sm__pipe_tensor') and h.startswith('sm__pipe_tensor') and 'cycles_active' in h and 'pct' in h):
            print('%-70s %-12s %s' % (h, units[i], [r[i] for r in rows[2:]]))
if src:
    rows = list(csv.reader(open(src)))
    hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
    ops, tot = collections.Counter(), 0
    st = collections.Counter()
    for r in rows[2:]:
        try: n = int(r[ix['Instructions Executed']])
        except Exception: continue
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
        op = m.group(2).split('.')[0] if m else '?'
        ops[op] += n; tot += n
        for h in hdr:
            if h.startswith('stall_') and 'Not Issued' not in h:
                try: st[h] += int(r[ix[h]])
                except Exception: pass
    print('warp instructions', tot)
    print('opcode mix:', ', '.join('%s %.1f%%' % (o, 100 * n / tot) for o, n in ops.most_common(14)))
    ts = sum(st.values()) or 1
    print('stall samples:', ', '.join('%s %.1f%%' % (k[6:], 100 * v / ts) for k, v in st.most_common(9)))
