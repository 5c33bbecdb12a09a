#!/usr/bin/env python
"""Developer aid for a container without a GPU: call the BODIES of the `-m gpu` tests with `DEV` pointed at the CPU and the
launch layer replaced by tests/sim_backend.py, to catch Python-level mistakes in a test (a wrong helper name, a shape, a
fixture key) before it meets the B200.  Tests that reach a kernel the stand-in does not model (direct C-ABI calls: the
Evaluator, the confidence scalars, training; the fused bf16 stem; CUDA graphs / pipelines) end in an AddError /
NotImplementedError / RuntimeError there — expected; exact batch-vs-single equalities (1e-6) may trip on oneDNN's
batch-dependent blocking.  Anything else is worth a look.  Proves nothing about the kernels.
Usage: python tools/gpu_tests_on_standin.py test_gpu_net test_gpu_ops test_zz_gpu_late ..."""
import sys, importlib, inspect, traceback, itertools
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'tests')); sys.path.insert(0, str(ROOT))
import pytest, torch
import sim_backend

class MP:
    def __init__(self): self.undo = []
    def setattr(self, obj, name, val):
        self.undo.append((obj, name, getattr(obj, name))); setattr(obj, name, val)
    def close(self):
        for obj, name, old in reversed(self.undo): setattr(obj, name, old)

def params_of(fn):
    marks = [m for m in getattr(fn, 'pytestmark', []) if m.name == 'parametrize']
    if not marks: return [{}]
    axes = []
    for m in marks:
        names = [n.strip() for n in m.args[0].split(',')] if isinstance(m.args[0], str) else list(m.args[0])
        vals = []
        for v in m.args[1]:
            v = getattr(v, 'values', v)
            if len(names) == 1 and not isinstance(v, (tuple, list)) : v = (v,)
            elif len(names) == 1 and isinstance(v,(tuple,list)) and len(v)!=1: v=(v,)
            vals.append(dict(zip(names, v)))
        axes.append(vals)
    out = []
    for combo in itertools.product(*axes):
        d = {}
        for c in combo: d.update(c)
        out.append(d)
    return out

def run(modname, only=None, skip=()):
    mod = importlib.import_module(modname)
    if hasattr(mod, 'DEV'): mod.DEV = torch.device('cpu') if isinstance(mod.DEV, torch.device) else 'cpu'
    ok = bad = 0
    for name, fn in inspect.getmembers(mod, inspect.isfunction):
        if not name.startswith('test_') or (only and name not in only) or name in skip: continue
        for kw in params_of(fn):
            sig = inspect.signature(fn)
            if any(p not in kw for p in sig.parameters):   # fixtures: skip
                print('SKIP (fixture)', name); break
            mp = MP(); sim_backend.install(mp)
            from add_b200.runtime import Plan
            mp.setattr(Plan, 'capture', lambda self, n_streams=None: None)
            try:
                fn(**kw); ok += 1
            except Exception as e:
                bad += 1; print('FAIL', modname, name, kw if len(str(kw))<120 else '', type(e).__name__, str(e)[:300])
                tb = traceback.extract_tb(e.__traceback__)
                print('   at', [(f.filename.split('/')[-1], f.lineno) for f in tb][-4:])
            finally:
                mp.close()
    print(modname, 'ok', ok, 'bad', bad)

if __name__ == '__main__':
    for m in sys.argv[1:]:
        run(m)
