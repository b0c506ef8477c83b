"""Read an `ncu --set full` report of scripts/ncu_roofline_targets.py here (no GPU needed) and write
profiles/ncu_traffic.json ({key: {bytes, source}}: what bench.py reports as `roofline.traffic`) plus a markdown summary.
usage: python scripts/ncu_traffic.py <report.ncu-rep> <celeba batch> <summary.md>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__m_xbar2l1tex_read_bytes.sum', 'lts__t_sector_hit_rate.pct',
           'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
           'sm__cycles_elapsed.avg.per_second', 'launch__registers_per_thread']


def to_bytes(v, unit):
    f = float(v.replace(',', ''))
    return f * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def main():
    rep, Bc, md = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = ['mnist_fashion_bf16_b1024_fprop', 'mnist_fashion_bf16_b1024_dgrad', 'celeba_bf16_b%d_conv7_fprop' % Bc,
            'celeba_bf16_b%d_conv7_dgrad' % Bc, 'celeba_bf16_b%d_conv7_wgrad' % Bc, 'celeba_bf16_b%d_in_style_resize' % Bc]
    out, lines = {}, ['# `ncu --set full` of the roofline launches (%s)\n' % os.path.basename(rep),
                      '| launch | kernel | ' + ' | '.join(m.split('.')[0] for m in METRICS) + ' | dram bytes (read + write) |', '|---|---|' + '---|' * (len(METRICS) + 1)]
    for key, r in zip(keys, rows[2:]):
        rd = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']])
        wr = to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
        out[key] = {'bytes': rd + wr, 'source': os.path.relpath(md, ROOT)}
        vals = ['%s %s' % (r[idx[m]], units[idx[m]]) if m in idx else '' for m in METRICS]
        lines.append('| %s | `%s` | ' % (key, r[idx['Kernel Name']][:60]) + ' | '.join(vals) + ' | %.1f MB |' % ((rd + wr) / 1e6))
    with open(md, 'w') as f:
        f.write('\n'.join(lines) + '\n')
    p = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    old = json.load(open(p)) if os.path.exists(p) else {}
    old.update(out)
    json.dump(old, open(p, 'w'), indent=1, sort_keys=True)
    print(open(md).read())


if __name__ == '__main__':
    main()
