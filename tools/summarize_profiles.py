"""Turns the ncu logs that tools/gpu_profile_final.sh leaves in gpurun_out/ into the compact evidence files under profiles/."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'profiles')
GO = os.path.join(ROOT, 'gpurun_out')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r02'


def read_ncu_csv(path):
    rows = []
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    for r in csv.DictReader(lines):
        rows.append(r)
    return rows


def short(name):
    for key in ('conv_pair_kernel', 'letterbox_s2d_kernel', 'nms_image_kernel', 'decode_kernel', 'upsample2x_kernel',
                'sppf_plane_kernel', 'sppf_pool_kernel', 'tta_scale_s2d_kernel', 'cand_from_pred_kernel', 'rotate_rgb_kernel'):
        if key in name:
            return key
    return name.split('(')[0][-60:]


def launches():
    p = os.path.join(GO, 'launches_bench.csv')
    if not os.path.exists(p):
        return
    rows = read_ncu_csv(p)
    agg = {}
    for r in rows:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', 'ns')
        ns = v * {'ns': 1, 'us': 1e3, 'ms': 1e6, 'nsecond': 1, 'usecond': 1e3, 'msecond': 1e6}.get(unit, 1)
        a = agg.setdefault(short(r['Kernel Name']), [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, TAG + '_launch_list_bench.csv'), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 3 (MDB_NO_GRAPHS=1)\n')
        f.write('# per-launch times under ncu are cold-cache and serialised: compare SHARES\n')
        f.write('kernel,launches,total_ms,share\n')
        for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('{},{},{:.3f},{:.4f}\n'.format(k, n, ns / 1e6, ns / tot))
    print('launch list:', {k: (v[0], round(v[1] / tot, 4)) for k, v in agg.items()})


def traffic():
    p = os.path.join(GO, 'traffic_step_bs64.csv')
    if not os.path.exists(p):
        return
    rows = read_ncu_csv(p)
    per = {}
    for r in rows:
        key = r['ID']
        d = per.setdefault(key, {'name': short(r['Kernel Name'])})
        v = float(r['Metric Value'].replace(',', ''))
        unit = r.get('Metric Unit', '')
        if r['Metric Name'].startswith('dram__bytes'):
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)
        d[r['Metric Name']] = v
    # keep exactly ONE forward pass: from the first letterbox launch to the first NMS launch
    ordered = [per[k] for k in sorted(per, key=lambda x: int(x))]
    first = next((i for i, d in enumerate(ordered) if d['name'] == 'letterbox_s2d_kernel'), 0)
    last = next((i for i, d in enumerate(ordered) if i >= first and d['name'] == 'nms_image_kernel'), len(ordered) - 1)
    ordered = ordered[first:last + 1]
    agg = {}
    for d in ordered:
        a = agg.setdefault(d['name'], [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get('dram__bytes_read.sum', 0.0)
        a[2] += d.get('dram__bytes_write.sum', 0.0)
    with open(os.path.join(OUT, TAG + '_dram_traffic_step_bs64.csv'), 'w') as f:
        f.write('# ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k <our kernels>  python tools/profile_layers.py 64 (MDB_NO_GRAPHS=1)\n')
        f.write('# = ONE eager forward pass over a batch of 64 images: K1, every conv launch, upsamples, SPPF pools, decode, NMS\n')
        f.write('kernel,launches,dram_read_GB,dram_write_GB\n')
        for k, (n, rd, wr) in sorted(agg.items(), key=lambda kv: -(kv[1][1] + kv[1][2])):
            f.write('{},{},{:.3f},{:.3f}\n'.format(k, n, rd / 1e9, wr / 1e9))
    print('traffic:', {k: (v[0], round((v[1] + v[2]) / 1e9, 2)) for k, v in agg.items()})


def reports():
    for f in sorted(os.listdir(GO)):
        if f.startswith('ncu_final_') and f.endswith('.ncu-rep'):
            base = f[:-8].replace('ncu_final_', TAG + '_ncu_')
            src = os.path.join(GO, f)
            raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
            open(os.path.join(OUT, base + '_raw.csv'), 'w').write(raw)
            det = subprocess.run(['ncu', '-i', src, '--page', 'details'], capture_output=True, text=True).stdout
            open(os.path.join(OUT, base + '_details.txt'), 'w').write('\n'.join(det.splitlines()[:220]) + '\n')
            print('report', base)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    launches()
    traffic()
    reports()
