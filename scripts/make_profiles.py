"""Summarise the ncu captures of one measurement pass into profiles/ (run in the build container).

usage: python scripts/make_profiles.py <tag> <fwd.ncu-rep> <bwd.ncu-rep> <launches.csv> <bench.json>
The captures come from (on the GPU box):
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
  ncu --set full --clock-control none --import-source on -k regex:lcp_forward  -s 1 -c 1 -o fwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline
  ncu --set full --clock-control none --import-source on -k regex:lcp_backward -s 1 -c 1 -o bwd python bench.py --steps 1 --warmup 1 --no-cpu-baseline
"""
import csv, json, shutil, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max',
        'smsp__cycles_active.avg', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
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
    cur, agg = None, []
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
        elif r[0].isdigit():
            try:
                s = int(r[4])
            except ValueError:
                continue
            det = sorted([(int(r[idx[h]]) if r[idx[h]].isdigit() else 0, h.replace('stall_', '')) for h in st], reverse=True)[:3]
            agg.append((s, cur, int(r[0]), r[1].strip()[:90], det))
    return agg


def main():
    tag, fwd, bwd, launches, bench = sys.argv[1:6]
    for name, rep in (('fwd', fwd), ('bwd', bwd)):
        d, u = raw(rep)
        with open('profiles/%s_%s_ncu_summary.txt' % (tag, name), 'w') as f:
            f.write('# ncu --set full --clock-control none --import-source on -k regex:lcp_%s -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline\n' % ('forward' if name == 'fwd' else 'backward'))
            f.write('# kernel: %s   (B = 4096 scenes, cfg3: n=96 m=256 fp32, one launch)\n' % d.get('Kernel Name'))
            for k in KEYS:
                if k in d:
                    f.write('%-70s %s %s\n' % (k, d[k], u.get(k, '')))
            st = {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): int(v) for h, v in d.items()
                  if h.startswith('smsp__pcsamp_warps_issue_stalled_') and 'not_issued' not in h}
            tot = sum(st.values()) or 1
            f.write('\n# warp stall sampling (all samples), share of %d samples\n' % tot)
            for k, v in sorted(st.items(), key=lambda x: -x[1]):
                if v:
                    f.write('%-24s %9d %5.1f%%\n' % (k, v, 100.0 * v / tot))
            agg = source_lines(rep)
            tot = sum(a[0] for a in agg) or 1
            f.write('\n# hottest source lines (stall samples, share, top-3 stall reasons)\n')
            for s, fl, l, sr, det in sorted(agg, reverse=True)[:45]:
                f.write('%8d %5.1f%% %s:%d  %s | %s\n' % (s, 100.0 * s / tot, fl, l, sr, ' '.join('%s:%d' % (h, v) for v, h in det)))
        tr = float(d['dram__bytes_read.sum']) * MULT[u['dram__bytes_read.sum']] + float(d['dram__bytes_write.sum']) * MULT[u['dram__bytes_write.sum']]
        json.dump({'kernel': d.get('Kernel Name'), 'batch': 4096, 'dram_bytes_per_launch': tr,
                   'duration_ms_under_ncu': float(d['gpu__time_duration.sum'])},
                  open('profiles/%s_%s_traffic.json' % (tag, name), 'w'))
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
