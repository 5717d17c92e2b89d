#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + sass pages) into text: key metrics, opcode mix, hottest SASS blocks."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
launch = sys.argv[2] if len(sys.argv) > 2 else '0'
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'local_load', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        # instruction supply: SM-level instruction cache (32 KiB, profiles/r2_icache_probe.jsonl) and the GPC-level
        # cache behind it, whose request rate is what bounded the round-1 interpreter and the 5-register-bit code
        'sm__icc_request_hit_rate.pct', 'sm__icc_requests.sum', 'gcc__cache_requests_type_instruction.sum',
        'gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed']
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print('%-88s %-14s %s' % (k, units[i], [r[i][:34] for r in rows[2:]]))
sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '--launch-skip', launch,
                       '--launch-count', '1'], capture_output=True, text=True).stdout
h = None
recs = []
for r in csv.reader(sass.splitlines()):
    if r and r[0] in ('Address', 'Line No'):
        h = r
        continue
    if h is None or len(r) < len(h) - 2:
        continue
    d = dict(zip(h, r))
    try:
        c = int(d['Instructions Executed'])
    except Exception:
        continue
    recs.append((c, d.get('Source', '')))
half = len(recs) // 2 if len(recs) > 2 and recs[:len(recs) // 2] == recs[len(recs) // 2:] else len(recs)
recs = recs[:half]
tot = sum(c for c, _ in recs)
ops = collections.Counter()
for c, s in recs:
    t = s.split()
    if not t:
        continue
    op = t[1] if t[0].startswith('@') else t[0]
    ops[op.split('.')[0]] += c
print('\nSASS instructions (static) %d, executed (warp-level) %.4e' % (len(recs), tot))
for k, v in ops.most_common(18):
    print('  %-10s %6.2f%%' % (k, 100.0 * v / tot))
blocks = []
cur = None
for c, s in recs:
    if cur and cur[0] == c:
        cur[1].append(s)
    else:
        cur = [c, [s]]
        blocks.append(cur)
print('\nhottest straight-line regions (exec count, length, share, mix)')
for b in sorted(blocks, key=lambda b: -b[0] * len(b[1]))[:16]:
    mix = collections.Counter((x.split()[1] if x.startswith('@') else x.split()[0]).split('.')[0] for x in b[1] if x.split())
    print('  %.3e x %4d = %5.2f%%  %s' % (b[0], len(b[1]), 100.0 * b[0] * len(b[1]) / tot, dict(mix.most_common(7))))
