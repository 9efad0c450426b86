"""Recipe for ``oracle/_ref`` — the REAL reference as a built artefact that can travel to the GPU box.

The reference is pure Python; "building" it means byte-compiling its own five ``core/`` modules from where they lie
under /root/reference (never copied as sources) into ``oracle/_ref/core/*.pyc.bin`` (bytecode under a neutral
extension: the gpurun snapshot drops ``*.pyc`` files).  ``oracle/_ref/`` is git-ignored (so
the history stays free of reference code) but not gpurun-ignored, so — like a compiled ``.so`` — it ships with the
snapshot to the GPU box, where /root/reference does not exist.

    python oracle/build_ref.py          # or __graft_entry__.build(), which calls build() when /root/reference exists

Used by ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference": the reference's own core/loss.py and
core/metric.py timed on the host cores), by ``bench.py``'s configs[1] leg (the reference's own DenseFuse) and by
``tests/`` to cross-check the oracle restatement against the real thing wherever the artefact is present.

TEST / BENCH INFRASTRUCTURE: nothing under multi-modal-image-fusion_b200/ may import this.
"""
import importlib.machinery
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = '/root/reference/core'
OUT = os.path.join(HERE, '_ref', 'core')
MODULES = ('fusion', 'block', 'model', 'loss', 'metric')       # import order: block needs fusion, model needs both
EXT = '.pyc.bin'                                                # python bytecode; '*.pyc' does not survive the gpurun snapshot
PKG = 'mmif_reference_core'                                     # private package name: never collides with dropin/core
# the reference's three scripts and what they import besides core/: byte-compiled too, so that the GPU box can run the
# UNMODIFIED train.py / test.py / eval.py with dropin/ first on sys.path (tests/test_reference_scripts_gpu.py)
REF_ROOT = '/root/reference'
SCRIPTS_OUT = os.path.join(HERE, '_ref', 'scripts')
SCRIPTS = ('train.py', 'test.py', 'eval.py', 'common.py', 'data/dataset.py', 'data/patches.py', 'data/transform.py')


def build(force=False):
    """Byte-compile the reference modules into oracle/_ref/core.  Returns True if the artefact is complete."""
    if not os.path.isdir(REF_SRC):
        return available()
    os.makedirs(OUT, exist_ok=True)
    for m in MODULES:
        src, dst = os.path.join(REF_SRC, m + '.py'), os.path.join(OUT, m + EXT)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f'reference/core/{m}.py', doraise=True, optimize=0,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    for rel in SCRIPTS:
        src, dst = os.path.join(REF_ROOT, rel), os.path.join(SCRIPTS_OUT, rel[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile=f'reference/{rel}', doraise=True, optimize=0,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(HERE, '_ref', 'BUILT_FROM'), 'w') as fh:
        fh.write(f'{REF_SRC} by oracle/build_ref.py with python {sys.version.split()[0]}\n')
    return available()


def available():
    return all(os.path.exists(os.path.join(OUT, m + EXT)) for m in MODULES)


def scripts_available():
    return available() and all(os.path.exists(os.path.join(SCRIPTS_OUT, rel[:-3] + EXT)) for rel in SCRIPTS)


def stage_scripts(dst_repo, with_reference_loss=False):
    """Lay the byte-compiled reference checkout out under `dst_repo` (train.pyc, test.pyc, eval.pyc, common.pyc, data/,
    core/{model,block,fusion}.pyc) the way the scripts expect to find their neighbours.  core/loss and core/metric are
    left out (they come from dropin/core, first on sys.path) unless `with_reference_loss` asks for the reference's own."""
    import shutil
    for rel in SCRIPTS:
        d = os.path.join(dst_repo, rel + 'c')
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(SCRIPTS_OUT, rel[:-3] + EXT), d)
    os.makedirs(os.path.join(dst_repo, 'core'), exist_ok=True)
    mods = MODULES if with_reference_loss else ('fusion', 'block', 'model')
    for m in mods:
        shutil.copyfile(os.path.join(OUT, m + EXT), os.path.join(dst_repo, 'core', m + '.pyc'))


def load():
    """Import the reference's modules from the artefact as the private package `mmif_reference_core` and return
    (loss, metric, model) modules.  Raises ImportError when oracle/_ref has not been built."""
    if PKG + '.loss' in sys.modules:
        return tuple(sys.modules[f'{PKG}.{m}'] for m in ('loss', 'metric', 'model'))
    if not available():
        raise ImportError('oracle/_ref is not built (run oracle/build_ref.py where /root/reference exists)')
    spec = importlib.machinery.ModuleSpec(PKG, None, is_package=True)
    pkg = importlib.util.module_from_spec(spec)
    pkg.__path__ = [OUT]
    sys.modules[PKG] = pkg
    for m in MODULES:
        name = f'{PKG}.{m}'
        loader = importlib.machinery.SourcelessFileLoader(name, os.path.join(OUT, m + EXT))
        mspec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(mspec)
        sys.modules[name] = mod
        loader.exec_module(mod)
        setattr(pkg, m, mod)
    return tuple(sys.modules[f'{PKG}.{m}'] for m in ('loss', 'metric', 'model'))


if __name__ == '__main__':
    ok = build(force=True)
    print('oracle/_ref', 'built' if ok else 'NOT available', OUT)
