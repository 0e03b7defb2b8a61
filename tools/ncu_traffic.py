"""Sum the DRAM traffic of the conv launches of one forward from an .ncu-rep (ncu --set full) -> profiles/r01_traffic_<mode>.json

usage: python tools/ncu_traffic.py rep.ncu-rep "<source note>" > profiles/r01_traffic_fp16x2.json
"""
import csv
import io
import json
import subprocess
import sys


def main(path, note):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    ni, ri, wi, ti = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum'), h.index('gpu__time_duration.sum')

    def gb(v, u):
        v = float(v.replace(',', ''))
        return v * {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}[u]

    def ms(v, u):
        v = float(v.replace(',', ''))
        return v * {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}[u]
    rd = wr = t = 0.0
    n = 0
    for r in rows[2:]:
        if not any(k in r[ni] for k in ('conv3d_tc_kernel', 'head_tc_kernel', 'conv3d_simt', 'conv3d_to1', 'space_to_depth')):
            continue
        rd += gb(r[ri], units[ri]); wr += gb(r[wi], units[wi]); t += ms(r[ti], units[ti]); n += 1
    print(json.dumps({'source': note, 'conv_launches': n, 'dram_read_gb_per_step': round(rd, 3), 'dram_write_gb_per_step': round(wr, 3),
                      'conv_ms_under_ncu': round(t, 3)}, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else '')
