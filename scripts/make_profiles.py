"""Summarise the ncu captures of one measurement pass into profiles/ (run in the build container).

usage: python scripts/make_profiles.py <tag> <launches.csv> <bench.json> <name>=<rep>:<batch>:<cfg> ...
e.g.   python scripts/make_profiles.py r02 gpurun_out/r02_launches_raw.csv gpurun_out/bench_r02b.json \
           fwd=gpurun_out/r02_fwd.ncu-rep:4096:cfg3 bwd=gpurun_out/r02_bwd.ncu-rep:4096:cfg3 \
           fwd_cfg2=gpurun_out/r02_fwd_cfg2.ncu-rep:1024:cfg2
The captures come from (on the GPU box):
  ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
  ncu --set full --metrics <local-memory / fp64-pipe counters> --clock-control none --import-source on -k regex:cond_forward -s 1 -c 1 -o fwd python scripts/prof_target.py 4096 cfg3
  (same with regex:cond_backward ... cfg3 bwd, and regex:cond_forward ... 1024 cfg2)
"""
import csv, json, shutil, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_pipe_fma.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']
MULT = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}


def raw(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    return dict(zip(rows[0], rows[-1])), dict(zip(rows[0], rows[1]))


def source_lines(rep):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = [r for r in rows if r and r[0] == 'Line No'][0]
    idx = {h: i for i, h in enumerate(hdr)}
    st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    cur, agg = None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
        elif r[0].isdigit() and len(r) > 6 and r[2] == '-':
            try:
                s, n = int(r[idx['# Samples']]), int(r[idx['Instructions Executed']])
            except ValueError:
                continue
            a = agg.setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:95], {}])
            a[0] += s; a[1] += n
            for h in st:
                v = r[idx[h]]
                if v.isdigit() and int(v):
                    a[3][h.replace('stall_', '')] = a[3].get(h.replace('stall_', ''), 0) + int(v)
    return agg


def main():
    tag, launches, bench = sys.argv[1:4]
    for spec in sys.argv[4:]:
        name, rest = spec.split('=')
        rep, batch, cfg = rest.split(':')
        d, u = raw(rep)
        with open('profiles/%s_%s_ncu_summary.txt' % (tag, name), 'w') as f:
            f.write('# ncu --set full (+ local-memory / fp64-pipe counters) --clock-control none --import-source on, one launch of\n')
            f.write('# kernel: %s   (B = %s scenes, %s; python scripts/%s)\n' % (d.get('Kernel Name'), batch, cfg, 'band_target.py' if cfg == 'cfg4' else 'prof_target.py'))
            for k in KEYS:
                if k in d:
                    f.write('%-72s %s %s\n' % (k, d[k], u.get(k, '')))
            st = {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): int(v) for h, v in d.items()
                  if h.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in h}
            tot = sum(st.values()) or 1
            f.write('\n# warp stall sampling (all samples), share of %d samples\n' % tot)
            for k, v in sorted(st.items(), key=lambda x: -x[1]):
                if v:
                    f.write('%-24s %9d %5.1f%%\n' % (k, v, 100.0 * v / tot))
            agg = source_lines(rep)
            tot = sum(a[0] for a in agg.values()) or 1
            ti = sum(a[1] for a in agg.values()) or 1
            f.write('\n# hottest source lines: stall samples, share | warp instructions executed, share | top-3 stall reasons\n')
            for (fl, l), (s, n, sr, det) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:50]:
                top = sorted(det.items(), key=lambda x: -x[1])[:3]
                f.write('%8d %5.1f%% | %10d %5.1f%% | %s:%d  %s | %s\n' % (s, 100.0 * s / tot, n, 100.0 * n / ti, fl, l, sr,
                                                                          ' '.join('%s:%d' % kv for kv in top)))
        tr = float(d['dram__bytes_read.sum']) * MULT[u['dram__bytes_read.sum']] + float(d['dram__bytes_write.sum']) * MULT[u['dram__bytes_write.sum']]
        suffix = '' if cfg == 'cfg3' else '_' + cfg
        json.dump({'kernel': d.get('Kernel Name'), 'batch': int(batch), 'dram_bytes_per_launch': tr,
                   'duration_ms_under_ncu': float(d['gpu__time_duration.sum']) * {'ms': 1, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}.get(u['gpu__time_duration.sum'], 1)},
                  open('profiles/%s_%s_traffic%s.json' % (tag, name.replace('_' + cfg, ''), suffix), 'w'))
        print(name, 'dram bytes/launch', tr, 'duration', d['gpu__time_duration.sum'], u['gpu__time_duration.sum'])
    rows = list(csv.reader(open(launches, errors='ignore')))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    out = [['id', 'kernel', 'stream', 'block', 'grid', 'gpu__time_duration.sum_ns']]
    for r in rows[hi + 1:]:
        if len(r) >= 15:
            out.append([r[0], r[4][:90], r[6], r[7], r[8], r[14]])
    with open('profiles/%s_launches.csv' % tag, 'w', newline='') as f:
        csv.writer(f).writerows(out)
    shutil.copy(bench, 'profiles/%s_bench_n1.json' % tag)


if __name__ == '__main__':
    main()
