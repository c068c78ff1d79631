#!/usr/bin/env python3
"""Turn the ncu artefacts a gpurun call left under gpurun_out/ into the tracked summaries under profiles/.

    ncu -i gpurun_out/prof_score_mma.ncu-rep --page raw --csv > /tmp/mma_raw.csv
    python tools/summarize_profiles.py score /tmp/mma_raw.csv profiles/r01_score_mma_kernel_ncu.md "<command>" "<title>"
    python tools/summarize_profiles.py launches gpurun_out/launches_bench.csv profiles/r01_bench_launches.md "<command>" "<title>"
"""
import collections
import csv
import json
import os
import re
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__ops_path_tensor_src_fp64.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second']


def score(src, dst, cmd, title, notes):
    rows = list(csv.reader(open(src)))
    d = {h: (rows[2][i], rows[1][i]) for i, h in enumerate(rows[0])}
    out = ['# ' + title, '', 'Command (under `gpurun`, after the same command exited 0 without ncu):', '', '```', cmd, '```', '',
           'Kernel: `%s`' % d['Kernel Name'][0], '', '| metric | value | unit |', '|---|---|---|']
    for k in KEYS:
        if k in d:
            out.append('| `%s` | %s | %s |' % (k, d[k][0], d[k][1]))
    st = sorted(((float(v[0]), h) for h, v in d.items() if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')), reverse=True)
    out += ['', 'Warp stall reasons (warps stalled per issue-active cycle, top 8):', '', '| reason | ratio |', '|---|---|']
    for v, h in st[:8]:
        out.append('| %s | %.2f |' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v))
    out += [''] + notes
    open(dst, 'w').write('\n'.join(out) + '\n')
    return (float(d['dram__bytes_read.sum'][0]) + float(d['dram__bytes_write.sum'][0])) * 1e6


def launches(src, dst, cmd, title, notes, skip=('peak_kernel',)):
    with open(src) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        name = re.sub(r'b200rbf::|\(anonymous namespace\)::|<unnamed>::|unnamed>::|void ', '', name).strip()
        if any(s in name for s in skip):
            continue
        v, u = float(row['Metric Value'].replace(',', '')), row['Metric Unit']
        v = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u in ('s', 'second') else v
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    out = ['# ' + title, '', '```', cmd, '```', '', 'Times are cold-cache and serialised under ncu: compare SHARES.  Total %.2f ms over %d launches.' % (T, sum(cnt.values())), '',
           '| kernel | launches | total ms | share | avg ms |', '|---|---|---|---|---|']
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        out.append('| `%s` | %d | %.3f | %.2f %% | %.4f |' % (k, cnt[k], v, 100 * v / T, v / cnt[k]))
    out += [''] + notes
    open(dst, 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
    kind, src, dst, cmd, title = sys.argv[1:6]
    notes = sys.argv[6:]
    if kind == 'score':
        t = score(src, dst, cmd, title, notes)
        print('dram bytes per launch: %.1f MB' % (t / 1e6))
    else:
        launches(src, dst, cmd, title, notes)
