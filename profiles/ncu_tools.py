"""Helpers to read ncu reports here (no GPU): raw metrics and hot SASS blocks.
usage: python profiles/ncu_tools.py raw  <rep> [kernel-regex]
       python profiles/ncu_tools.py hot  <rep> <kernel-regex>
       python profiles/ncu_tools.py launches <csv>"""
import collections
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio']


def ncu(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def raw(rep, kre=None):
    a = ['-i', rep, '--page', 'raw', '--csv']
    if kre:
        a += ['--kernel-name', 'regex:' + kre]
    rows = list(csv.reader(io.StringIO(ncu(a))))
    h = rows[0]
    for r in rows[2:]:
        print('-----', r[h.index('Kernel Name')][:70])
        for w in WANT:
            if w in h:
                print('  %-86s %s %s' % (w, r[h.index(w)], rows[1][h.index(w)]))


def hot(rep, kre, thresh=0.01):
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre]))))
    h = rows[1]
    ia, isrc, ismp = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples')
    data = [(int(r[ia]), int(r[ismp]), r[isrc].strip(), i) for i, r in enumerate(rows[2:]) if len(r) > ia and r[ia].isdigit()]
    tot = sum(d[0] for d in data)
    tots = sum(d[1] for d in data)
    print('total inst', tot, 'sass lines', len(data), 'samples', tots)
    prev, start, acc, accs, blocks = None, 0, 0, 0, []
    for d in data:
        key = d[0]
        if prev is None or abs(key - prev) > 0.02 * max(key, prev, 1):
            if prev is not None:
                blocks.append((start, d[3] - 1, prev, acc, accs))
            start, acc, accs, prev = d[3], 0, 0, key
        acc += d[0]
        accs += d[1]
    blocks.append((start, data[-1][3], prev, acc, accs))
    for b in blocks:
        if b[3] > thresh * tot or b[4] > thresh * tots:
            print('sass %4d-%4d exec/inst=%10d n=%3d inst_share=%.3f sample_share=%.3f first: %s'
                  % (b[0], b[1], b[2], b[1] - b[0] + 1, b[3] / tot, b[4] / max(tots, 1), data[b[0]][2][:50]))
    return data


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    h = rows[hi]
    kn, mv = h.index('Kernel Name'), h.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split('(')[0][:60]
        agg.setdefault(name, [0, 0.0])
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(',', ''))
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-62s n=%4d total=%10.1f us avg=%9.1f us share=%.3f' % (k, v[0], v[1] / 1e3, v[1] / v[0] / 1e3, v[1] / tot))


if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'raw':
        raw(*sys.argv[2:])
    elif cmd == 'hot':
        hot(sys.argv[2], sys.argv[3])
    elif cmd == 'launches':
        launches(sys.argv[2])
