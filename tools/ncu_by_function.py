"""Developer tool: attribute the warp-stall samples and executed instructions of an `ncu --set full --import-source on` capture
of k_step to the functions of rs_core.h / rs_env.h / rs_api.cu.

    python tools/ncu_by_function.py gpurun_out/<capture>.ncu-rep build/variants/<library built with -lineinfo>.so [kernel-substring]

ncu's CSV source page lists SASS only; the line table comes from `nvdisasm -g` on the cubin of the same library (its
`//## File "...", line N inlined at ...` annotations), joined on the instruction offset.  Each instruction is charged to the
function that holds its source line (innermost frame) and, in the second table, to the phase (outermost rs_core.h function)."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, lib = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else 'k_stepILi4ELi4'
LINES_OF = sys.argv[4] if len(sys.argv) > 4 else None      # optional: per-source-line table of this function (its own frame in the inline chain)
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def function_ranges(path):
    """(first line, last line, name) of every top-level function in a header, by brace matching."""
    out, depth, cur, name = [], 0, None, None
    pat = re.compile(r'^(?:RS_HD|__device__|__global__|static|template|inline)?.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;]*$')
    lines = open(path).read().split('\n')
    pending = None
    for i, ln in enumerate(lines, 1):
        code = ln.split('//')[0]
        if depth <= 1 and cur is None:
            m = re.match(r'^\s*(?:template\s*<[^>]*>\s*)?(?:RS_HD|__device__ __forceinline__|__device__|__global__|static|inline)[^;=]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(', code)
            if m and not code.strip().startswith('#'):
                pending = (i, m.group(1))
        opens, closes = code.count('{'), code.count('}')
        if pending and opens and cur is None:
            cur, name, base = pending[0], pending[1], depth
            pending = None
        depth += opens - closes
        if cur is not None and depth <= base:
            out.append((cur, i, name)); cur = None
    return out


srcs = {}
for f in ('robosumo_selfplay_b200/csrc/rs_core.h', 'robosumo_selfplay_b200/csrc/rs_env.h', 'robosumo_selfplay_b200/csrc/rs_api.cu'):
    srcs[os.path.basename(f)] = function_ranges(os.path.join(ROOT, f))


def fn_of(fname, line):
    for a, b, n in srcs.get(os.path.basename(fname), []):
        if a <= line <= b:
            return n
    return os.path.basename(fname) + ':outer'


tmp = tempfile.mkdtemp()
subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '-gi', '-c', cubin], capture_output=True, text=True).stdout.split('\n')
start = next(i for i, l in enumerate(dis) if l.startswith('.text.') and kern in l)
chain_at = {}
cur, fresh = [], True          # consecutive annotation lines = the inline chain, innermost frame first
for l in dis[start + 1:]:
    if l.startswith('//----') or (l.startswith('.text.') and kern not in l):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            cur, fresh = [], False
        cur.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]+)\*/\s+(\S.*);', l)
    if m:
        chain_at[int(m.group(1), 16)] = cur
        fresh = True

raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[h]
col = {n: i for i, n in enumerate(hdr)}
body = [r for r in rows[h + 1:] if len(r) == len(hdr)]
base = int(body[0][0], 16)
stalls = [n for n in hdr if n.startswith('stall_') and '(Not Issued)' not in n]
agg_in, agg_out = defaultdict(lambda: defaultdict(float)), defaultdict(lambda: defaultdict(float))
by_line = defaultdict(lambda: defaultdict(float))
tot = defaultdict(float)
for r in body:
    off = int(r[0], 16) - base
    ch = chain_at.get(off, [])
    inner = fn_of(*ch[0]) if ch else '?'
    # phase = first frame below the drivers (kernel body, simulate, eval_begin, solve_iter, ...), walking from the outside in
    names = [fn_of(f, ln) for f, ln in ch][::-1]
    drivers = ('__launch_bounds__', 'simulate', 'simulate_trips', 'eval_begin', 'forward', 'solve', 'solve_iter', 'rs_api.cu:outer')
    outer = next((n for n in names if n not in drivers), names[-1] if names else '?')
    if outer != '?' and 'solve_iter' in names and outer not in ('jt_forces', 'build_H_arrow', 'build_H', 'arrow_solve', 'chol_solve', 'gj_rows_in_registers', 'line_search', 'woodbury_solve', 'arrow_solve_multi'):
        outer = 'iter:' + outer if outer in ('twists', 'rows_of', 'mat_vec') else ('solve_iter' if names[-1] == 'solve_iter' or outer in ('ld3', 'v3', 'dot', 'cross') else outer)
    if 'solve_first' in names:
        outer = 'solve_first'
    ex = float(r[col['Instructions Executed']] or 0); thr = float(r[col['Thread Instructions Executed']] or 0)
    smp = float(r[col['# Samples']] or 0)
    if LINES_OF:
        for f, ln in ch:
            if fn_of(f, ln) == LINES_OF:
                by_line[ln]['exec'] += ex; by_line[ln]['thr'] += thr; by_line[ln]['sass'] += 1; by_line[ln]['samples'] += smp - float(r[col['stall_barrier']] or 0)
                break
    for A, key in ((agg_in, inner), (agg_out, outer)):
        A[key]['exec'] += ex; A[key]['thr'] += thr; A[key]['samples'] += smp; A[key]['sass'] += 1
        for s in stalls:
            A[key][s] += float(r[col[s]] or 0)
    tot['exec'] += ex; tot['samples'] += smp
    for s in stalls:
        tot[s] += float(r[col[s]] or 0)
print('total: warp-instructions %.0f, samples %.0f; stall samples: %s' %
      (tot['exec'], tot['samples'], ', '.join('%s %.1f%%' % (s[6:], 100 * tot[s] / tot['samples']) for s in sorted(stalls, key=lambda s: -tot[s])[:8])))
for title, A in (('by innermost function', agg_in), ('by phase (outermost function below the kernel body)', agg_out)):
    print('\n' + title)
    print('%-26s %7s %7s %7s %6s %6s | non-barrier stall samples: %s' % ('function', 'exec%', 'smp%', 'nonbar%', 'lanes', 'sass', 'wait short_sb not_sel no_inst branch mio'))
    nb_tot = tot['samples'] - tot['stall_barrier']
    for k, v in sorted(A.items(), key=lambda kv: -kv[1]['exec'])[:28]:
        nb = v['samples'] - v['stall_barrier']
        print('%-26s %7.2f %7.2f %7.2f %6.1f %6d | %5.1f %5.1f %5.1f %5.1f %5.1f %5.1f' % (
            k, 100 * v['exec'] / tot['exec'], 100 * v['samples'] / tot['samples'], 100 * nb / nb_tot, v['thr'] / max(v['exec'], 1), v['sass'],
            100 * v['stall_wait'] / max(nb, 1), 100 * v['stall_short_sb'] / max(nb, 1), 100 * v['stall_not_selected'] / max(nb, 1),
            100 * v['stall_no_inst'] / max(nb, 1), 100 * v['stall_branch_resolving'] / max(nb, 1), 100 * v['stall_mio'] / max(nb, 1)))

if LINES_OF:
    src = open(os.path.join(ROOT, 'robosumo_selfplay_b200/csrc/rs_core.h')).read().split('\n')
    print('\nlines of %s (warp-instructions executed per pair-step assuming 4096 pairs, lanes, distinct SASS, non-barrier samples)' % LINES_OF)
    for ln, v in sorted(by_line.items()):
        print('%5d %8.0f %5.1f %5d %6.0f | %s' % (ln, v['exec'] / 4096, v['thr'] / max(v['exec'], 1), v['sass'], v['samples'], src[ln - 1].strip()[:110]))
