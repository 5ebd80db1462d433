"""Bring-up tool: run every GPU parity check in its own process (a faulting kernel poisons the
CUDA context) with a timeout, and write a summary to gpurun_out/diag.json.

    python tests/gpu_diag.py                 # all checks
    python tests/gpu_diag.py NAME [NAME...]  # selected checks (in-process if exactly one + --one)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_one(name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import gpu_checks
    res = gpu_checks.CHECKS[name]()
    import torch
    torch.cuda.synchronize()
    print('RESULT ' + json.dumps(res))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    if '--one' in sys.argv:
        return run_one(args[0])
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    # names only (do not initialise CUDA in the parent)
    import ast
    src = open(os.path.join(ROOT, 'tests', 'gpu_checks.py')).read()
    names = []
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.Assign) and getattr(node.targets[0], 'id', None) == 'CHECKS':
            names = [k.value for k in node.value.keys]
    if args:
        names = [n for n in names if any(a in n for a in args)]
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    summary = {}
    env = dict(os.environ)
    env['PYTHONUNBUFFERED'] = '1'
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), '--one', name],
                               capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
            out = p.stdout + p.stderr
            ok = p.returncode == 0
            res = None
            for line in p.stdout.splitlines():
                if line.startswith('RESULT '):
                    res = json.loads(line[7:])
            tail = '' if ok else out[-1500:]
        except subprocess.TimeoutExpired as e:
            ok, res, tail = False, None, 'TIMEOUT ' + str(e)[-300:]
        summary[name] = {'ok': ok, 'res': res, 'sec': round(time.time() - t0, 1), 'tail': tail}
        print(('PASS ' if ok else 'FAIL ') + name, json.dumps(res), f'({summary[name]["sec"]}s)', flush=True)
        if not ok:
            print('    ' + tail.replace('\n', '\n    '), flush=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'diag.json'), 'w') as f:
            json.dump(summary, f, indent=1)
    n_ok = sum(1 for v in summary.values() if v['ok'])
    print(f'{n_ok}/{len(summary)} checks passed')


if __name__ == '__main__':
    main()
