"""Builds lightcurver_b200/liblcb.so (sm_100a only) with nvcc, in-tree.

``python -m lightcurver_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a
GPU; the resulting .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / 'csrc'
OUT = HERE / 'liblcb.so'
OBJ = HERE / 'build'

NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '--extended-lambda']


def _needs(target: Path, deps):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(d).stat().st_mtime > t for d in deps)


def build(verbose=False, force=False, ptxas_info=False, extra_flags=(), out=None):
    """extra_flags/out: build a variant (e.g. ['-DLCB_PHASE_TIMERS'] -> liblcb_timers.so for tools/phase_time.py)."""
    global OUT
    if out is not None or extra_flags:
        force = True
    OBJ.mkdir(exist_ok=True)
    sources = sorted(CSRC.glob('*.cu'))
    headers = sorted(CSRC.glob('*.cuh')) + [HERE.parent / 'include' / 'lcb.h']
    jobs = []
    for src in sources:
        obj = OBJ / (src.stem + '.o')
        if force or _needs(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + list(extra_flags) + (['-Xptxas', '-v'] if ptxas_info else []) + ['-c', str(src), '-o', str(obj)]
            jobs.append((src, cmd))
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = {ex.submit(subprocess.run, cmd, capture_output=True, text=True): src for src, cmd in jobs}
            for fut in concurrent.futures.as_completed(futs):
                r = fut.result()
                if verbose or r.returncode != 0 or ptxas_info:
                    sys.stderr.write(r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError(f'nvcc failed on {futs[fut]}')
    objs = [str(OBJ / (s.stem + '.o')) for s in sources]
    target = Path(out) if out is not None else OUT
    if jobs or not target.exists():
        cmd = [NVCC, '-shared', '-o', str(target)] + objs + ['-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    if out is not None:          # variant objects must not be mistaken for the default build
        for o in objs:
            Path(o).unlink(missing_ok=True)
    return target


if __name__ == '__main__':
    if '--timers' in sys.argv:
        p = build(extra_flags=['-DLCB_PHASE_TIMERS'], out=HERE / 'liblcb_timers.so')
        build(force=True)
    elif '--variant' in sys.argv:
        # python -m lightcurver_b200.build --variant NAME -DFLAG ... -> liblcb_NAME.so next to the default library (A/B timing)
        i = sys.argv.index('--variant')
        p = build(extra_flags=[a for a in sys.argv[i + 2:] if a.startswith('-')], out=HERE / f'liblcb_{sys.argv[i + 1]}.so')
        build(force=True)
    else:
        p = build(verbose='-v' in sys.argv, force='-f' in sys.argv, ptxas_info='--ptxas' in sys.argv)
    print(p)
