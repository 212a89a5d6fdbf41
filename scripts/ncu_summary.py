"""Build container (no GPU): summarise the `ncu --set full` captures of the decode kernel brought back in gpurun_out/ into profiles/.
Usage: ncu_summary.py  -> profiles/r02_ncu_decode_by_length.json, profiles/r02_ncu_decode_L*_details.txt, profiles/decode_traffic.json"""
import csv
import io
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, KV = 1361504256, 147456
PEAK = json.load(open(os.path.join(REPO, 'MEASURED_PEAKS.json')))['hbm_gbs'] if os.path.exists(os.path.join(REPO, 'MEASURED_PEAKS.json')) else 6650.0


def raw_metrics(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(m, key):
    v, u = m[key]
    x = float(v.replace(',', ''))
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'msecond': 1e-3, 'usecond': 1e-6, 'second': 1.0, 'nsecond': 1e-9}.get(u, 1.0)
    return x * scale


def main():
    res = {'note': 'ncu --set full --clock-control none, one launch of 8 sampled tokens (7 forward passes) at three cache lengths; times under the '
                   'profiler (serialised, cold) are NOT bench values', 'kernel': 'er::decode_persistent_kernel<false,true> (tensor-parallel layer)', 'captures': {}}
    for name, L0 in (('L2050', 2050), ('L8050', 8050), ('L17950', 17950)):
        rep = os.path.join(REPO, 'gpurun_out', f'r02_decode_{name}.ncu-rep')
        if not os.path.exists(rep):
            continue
        m = raw_metrics(rep)
        rd, wr, t = num(m, 'dram__bytes_read.sum'), num(m, 'dram__bytes_write.sum'), num(m, 'gpu__time_duration.sum')
        n = 7
        alg = n * (W + KV) + KV * (n * L0 + n * (n - 1) // 2)
        res['captures'][name] = {
            'cache_rows_first_pass': L0, 'forward_passes': n, 'gpu_time_ms': t * 1e3, 'dram_read_GB': rd / 1e9, 'dram_write_MB': wr / 1e6,
            'algorithmic_GB': alg / 1e9, 'dram_over_algorithmic': (rd + wr) / alg, 'dram_GBps_under_ncu': (rd + wr) / t / 1e9,
            'frac_of_measured_peak_under_ncu': (rd + wr) / t / 1e9 / PEAK, 'tokens_per_s_under_ncu': n / t,
            'registers_per_thread': m.get('launch__registers_per_thread', ('', ''))[0], 'shared_mem_per_block': m.get('launch__shared_mem_per_block_dynamic', ('', ''))[0]}
        det = subprocess.run(['ncu', '-i', rep, '--page', 'details'], capture_output=True, text=True).stdout
        open(os.path.join(REPO, 'profiles', f'r02_ncu_decode_{name}_details.txt'), 'w').write(det)
        if name == 'L8050':
            json.dump({'kernel': res['kernel'], 'capture': f'ncu --set full --clock-control none --import-source on -k regex:decode_persistent -s 1 -c 1 python scripts/ncu_decode.py 8 6000 (profiles/r02_ncu_decode_L8050_details.txt)',
                       'launch': '8 sampled tokens = 7 forward passes at cache lengths 8050..8056 (after an un-profiled 6000-token launch)',
                       'dram_bytes_read': rd, 'dram_bytes_write': wr, 'dram_bytes_per_launch': rd + wr, 'algorithmic_bytes_per_launch': alg,
                       'dram_over_algorithmic': (rd + wr) / alg, 'gpu_time_duration_ms': t * 1e3,
                       'note': 'traffic = ncu dram__bytes_read.sum + dram__bytes_write.sum of a 7-pass launch at L~8050 scaled by algorithmic bytes: ncu cannot replay the 12 s bench launch'},
                      open(os.path.join(REPO, 'profiles', 'decode_traffic.json'), 'w'), indent=1)
    json.dump(res, open(os.path.join(REPO, 'profiles', 'r02_ncu_decode_by_length.json'), 'w'), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
