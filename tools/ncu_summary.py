"""gpurun_out/prof_<k>_raw.csv (ncu --page raw --csv of one --set full capture) -> profiles/<prefix>_<k>_ncu.txt: the handful of
metrics DESIGN.md / the bench line quote, plus the SASS instruction census of the shipped library (cuobjdump)."""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
METRICS = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
           'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
           'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
           'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
           'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
           'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'launch__registers_per_thread',
           'launch__shared_mem_per_block_dynamic', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
           'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']


def summarise(kind, prefix, cmd):
    src = os.path.join(ROOT, 'gpurun_out', 'prof_%s_raw.csv' % kind)
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    d = {h: (v, u) for h, v, u in zip(rows[0], rows[2], rows[1])}
    out = os.path.join(ROOT, 'profiles', '%s_%s_ncu.txt' % (prefix, kind))
    with open(out, 'w') as f:
        f.write('# ncu --set full --clock-control none, one launch of %s\n# command: %s\n' % (d.get('Kernel Name', ('?',))[0], cmd))
        f.write('# (numbers under a profiler: cold caches, serialised -- shares and ratios, not bench values)\n')
        for m in METRICS:
            if m in d:
                f.write('%-82s %s %s\n' % (m, d[m][0], d[m][1]))
    print('wrote', out)


def summarise_many(kind, prefix, cmd):
    """Same, for a capture holding several kernels (one block per launch)."""
    src = os.path.join(ROOT, 'gpurun_out', 'prof_%s_raw.csv' % kind)
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    out = os.path.join(ROOT, 'profiles', '%s_%s_ncu.txt' % (prefix, kind))
    with open(out, 'w') as f:
        f.write('# ncu --set full --clock-control none\n# command: %s\n' % cmd)
        f.write('# (numbers under a profiler: cold caches, serialised -- shares and ratios, not bench values)\n')
        for r in rows[2:]:
            d = {h: (v, u) for h, v, u in zip(rows[0], r, rows[1])}
            f.write('\n## %s\n' % d.get('Kernel Name', ('?',))[0][:160])
            for m in METRICS:
                if m in d:
                    f.write('%-82s %s %s\n' % (m, d[m][0], d[m][1]))
    print('wrote', out)


def sass_census(prefix):
    lib = os.path.join(ROOT, 'scenedreamer_b200', 'libsdb200.so')
    txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    per, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            per[cur] = {}
            continue
        if cur:
            for op in ('UTCHMMA', 'UTCQMMA', 'LDTM', 'UBLKCP', 'UTCBAR', 'UTMALDG', 'HMMA', 'SYNCS', 'REDG', 'RED.E', 'ATOMG'):
                if re.search(r'\b%s' % re.escape(op), line):
                    per[cur][op] = per[cur].get(op, 0) + 1
    out = os.path.join(ROOT, 'profiles', '%s_sass_census.txt' % prefix)
    with open(out, 'w') as f:
        f.write('# cuobjdump -sass scenedreamer_b200/libsdb200.so: tensor-core / TMA / TMEM instruction counts per kernel (sm_100a)\n')
        f.write('# UTCHMMA = tcgen05.mma (kind::f16), LDTM = tcgen05.ld (TMEM -> registers), UBLKCP = cp.async.bulk (1-D TMA),\n')
        f.write('# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA (legacy mma.sync) must be absent.\n')
        tot = {}
        for fn, c in sorted(per.items()):
            if c:
                name = subprocess.run(['c++filt', fn], capture_output=True, text=True).stdout.strip()[:150]
                f.write('%-150s %s\n' % (name, ' '.join('%s=%d' % kv for kv in sorted(c.items()))))
                for k, v in c.items():
                    tot[k] = tot.get(k, 0) + v
        f.write('TOTAL %s\n' % ' '.join('%s=%d' % kv for kv in sorted(tot.items())))
    print('wrote', out, tot)


if __name__ == '__main__':
    prefix = sys.argv[1] if len(sys.argv) > 1 else 'r02'
    summarise('render', prefix, 'python bench.py --steps 2 --warmup 3 --no-cpu --no-extras  (-k regex:mlp_kernel -s 9 -c 1; the ray-slot kernel mlp_kernel<2,0,0,0,1>)')
    summarise('conv', prefix, 'python bench.py ...  (-k regex:conv_kernel -s 8 -c 1: conv2a, 3x3 256->256, fp16x3)')
    summarise('dda', prefix, 'python bench.py ...  (-k regex:dda_perspective -s 3 -c 1)')
    summarise('wgrad', prefix, 'python bench_train.py --steps 2 --warmup 3 --no-composition  (-k regex:wgrad_kernel -s 2 -c 1)')
    summarise_many('gridenc', prefix, 'python tools/gridenc_run.py  (-k regex:grid_|input_backward -s 3 -c 3: forward, table backward, input backward; 599k samples, D=5)')
    sass_census(prefix)
