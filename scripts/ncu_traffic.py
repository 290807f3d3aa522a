"""profiles/ncu_traffic.json + a text summary from ONE `ncu --set full` capture of scripts/ncu_r2_target.py.

    python scripts/ncu_traffic.py gpurun_out/ncu_r2.ncu-rep r2a        (runs here: `ncu -i` needs no GPU)

For every named launch (scripts/ncu_r2_target.py::ORDER; the second launch of each pair): dram__bytes_read.sum + dram__bytes_write.sum,
duration, tensor-pipe and DRAM utilisation, shared-memory bank conflicts.  bench.py reads the JSON for `roofline.traffic`."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scripts.ncu_r2_target import ORDER          # noqa: E402

WANT = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__cluster_size']


def main(rep, tag):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    name_i = hdr.index('Kernel Name')
    launches = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        rec = {'kernel': r[name_i]}
        for i, h in enumerate(hdr):
            if h in WANT or any(h.startswith(w) for w in WANT):
                try:
                    rec[h] = float(r[i].replace(',', ''))
                    rec[h + ':unit'] = units[i]
                except ValueError:
                    pass
        launches.append(rec)
    picked = launches[1::2]                     # second launch of each pair
    assert len(picked) >= len(ORDER), (len(launches), 'launches captured; expected', 2 * len(ORDER))
    traffic = {'source': f'{os.path.basename(rep)} ({tag}); ncu --set full --clock-control none, second launch of each pair in scripts/ncu_r2_target.py'}
    lines = [f'ncu --set full capture {tag}: one launch per hot kernel at BASELINE configs[1] shapes (N = 32)', '']
    for name, rec in zip(ORDER, picked):
        def val(key):
            for k, v in rec.items():
                if k == key:
                    u = rec.get(k + ':unit', '')
                    scale = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'byte': 1.0, 'usecond': 1e-3, 'us': 1e-3, 'msecond': 1.0, 'ms': 1.0, 'nsecond': 1e-6, 'ns': 1e-6, 'second': 1e3, 's': 1e3}.get(u, 1.0)
                    return v * scale
            return None
        rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
        traffic[name] = int(rd + wr) if rd is not None and wr is not None else None
        lines.append(f'{name:20s} {rec["kernel"][:70]}')
        lines.append(f'    duration {val("gpu__time_duration.sum")} ms   DRAM read {rd / 1e6 if rd else None} MB + write {wr / 1e6 if wr else None} MB')
        for k in WANT[3:]:
            if k in rec:
                lines.append(f'    {k} = {rec[k]} {rec.get(k + ":unit", "")}')
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    json.dump(traffic, open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json'), 'w'), indent=1)
    open(os.path.join(ROOT, 'profiles', f'ncu_{tag}_summary.txt'), 'w').write('\n'.join(lines) + '\n')
    print('\n'.join(lines))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'r2')
