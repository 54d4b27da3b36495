"""Per-phase instruction / stall breakdown of a k_emit ncu capture: python scripts/emit_phases.py rep [tile_pairs]"""
import collections, csv, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]; tp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_lines.py'), rep, 'k_emitILi2', '2000'], capture_output=True, text=True).stdout.split('\n')
tot_inst = int(re.search(r'instructions executed: (\d+)', out[0]).group(1))
src = open(os.path.join(ROOT, 'mimick_b200/csrc/mmk_emit.cuh')).read().split('\n')
def find(pat): return next(i + 1 for i, l in enumerate(src) if pat in l)
marks = [('decode_pair', find('__device__ __forceinline__ void decode_pair')), ('generic', find('// 0-based haplotype position of base j')),
         ('tile_prepare', find('k_tile_prepare(')), ('expand16', find('__device__ __forceinline__ uint4 expand16')),
         ('sts helpers', find('// shared memory through 32-bit window addresses')), ('fastdesc', find('// Per-pair parameters of the fast path')),
         ('phase1 seq', find('__device__ __forceinline__ void emit_pair_seq')), ('phase2 hdr', find('__device__ __forceinline__ void emit_header_piece')),
         ('copy_out', find('// stage → global')), ('kernel body', find('k_emit(const __grid_constant__ EmitArgs a)')), ('end', len(src) + 1)]
agg, stl = collections.Counter(), collections.Counter()
for l in out[1:]:
    m = re.match(r'\s*([\d.]+)% inst\s+([\d.]+)% stall-samples\s+(\S+):(\d+)', l)
    if not m: continue
    pi, ps, f, ln = float(m.group(1)), float(m.group(2)), m.group(3), int(m.group(4))
    g = f
    if f == 'mmk_emit.cuh':
        for (n, a), (n2, b) in zip(marks, marks[1:]):
            if a <= ln < b: g = n; break
    agg[g] += pi; stl[g] += ps
rows = list(csv.reader(subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout.split('\n')))
h, r = rows[0], rows[2]
pairs = int(r[h.index('launch__grid_size')]) * tp
for g in agg: print(f"{g:20s} {agg[g]:5.1f}% inst = {agg[g] / 100 * tot_inst / pairs:6.1f} instr/pair  {stl[g]:5.1f}% stalls")
print('total/pair', round(tot_inst / pairs, 1))
items = [(k, r[i]) for i, k in enumerate(h) if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k]
tot = sum(float(v.replace(',', '')) for k, v in items)
for k, v in sorted(items, key=lambda x: -float(x[1].replace(',', '')))[:7]: print(f'{k:60s} {100 * float(v.replace(",", "")) / tot:5.1f}%')
for k in ['gpu__time_duration.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
          'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_write.sum', 'dram__bytes_read.sum', 'launch__registers_per_thread',
          'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps']:
    if k in h: print(k, r[h.index(k)])
