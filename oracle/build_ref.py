"""Compile the UNMODIFIED reference into oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.build_ref          # /root/reference  ->  oracle/_ref/t2onet/**.t2oc (+ the vocabulary JSONs)

The reference is pure Python, so "building" it means byte-compiling its modules where they lie under
/root/reference (compile() + marshal, nothing is edited) into a sourceless tree: oracle/_ref/t2onet/models/actor.t2oc,
executors/executor.t2oc, utils/beam_search.t2oc, ... (the extension is not .pyc because the GPU-box snapshot drops
*.pyc files; oracle/ref_shims.py installs the import hook that loads them).  No reference SOURCE enters the repository: oracle/_ref/ is
git-ignored (not gpurun-ignored), so the compiled files travel to the GPU box next to the repo's own built .so and
let the `-m gpu` tests run the reference's own Actor / Executor / beam_search there (tests/test_gpu_actor.py) and
`bench.py --impl reference` time the reference itself (cpu_baseline.kind = "reference").  The sourceless tree is
imported exactly like the source tree, through oracle/ref_shims.py (same import shims).

The byte-code is tied to this image's Python (3.12); the GPU box runs the same image.
"""
import importlib.util
import marshal
import os
import shutil
import sys

SRC = os.environ.get('T2O_REFERENCE_ROOT', '/root/reference')
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref', 't2onet')
# the path's modules and their import closure (the actor and its encoders / decoder, the executor and operators,
# the planners and the options they are constructed from); nothing of pyutils/ (EdgeConnect) or the GAN variants
PY_DIRS = ['models', 'executors', 'utils', 'utils/ssim', 'options', 'datasets']
EXT = '.t2oc'          # importlib MAGIC_NUMBER + marshal.dumps(code object)
DATA = ['data/language/FiveK_operator_vocabs_sess_1.json', 'data/language/FiveK_vocabs_sess_1.json',
        'data/language/GIER_operator_vocabs_sess_3.json', 'data/language/GIER_vocabs_sess_3.json']


def available():
    return os.path.isfile(os.path.join(SRC, 'models', 'operators.py'))


def build(verbose=True):
    if not available():
        if verbose:
            print('[oracle.build_ref] %s not present: keeping the prebuilt oracle/_ref (if any)' % SRC)
        return None
    n = 0
    shutil.rmtree(DST, ignore_errors=True)
    for d in PY_DIRS:
        sdir = os.path.join(SRC, d)
        if not os.path.isdir(sdir):
            continue
        for f in sorted(os.listdir(sdir)):
            if not f.endswith('.py'):
                continue
            out = os.path.join(DST, d, f[:-3] + EXT)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            try:
                with open(os.path.join(sdir, f), 'rb') as fh:
                    code = compile(fh.read(), os.path.join('reference', d, f), 'exec', dont_inherit=True)
                with open(out, 'wb') as fh:
                    fh.write(importlib.util.MAGIC_NUMBER + marshal.dumps(code))
                n += 1
            except SyntaxError as e:                        # a module of the reference that does not parse on 3.12
                if verbose:
                    print('[oracle.build_ref] skipped %s/%s: %s' % (d, f, e))
    for f in DATA:
        out = os.path.join(DST, f)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), out)
    if verbose:
        print('[oracle.build_ref] %d modules -> %s' % (n, DST))
    return DST


if __name__ == '__main__':
    sys.exit(0 if build() or not available() else 1)
