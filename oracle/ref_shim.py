"""Import the REAL reference ``/root/reference/karma/kmer.py`` (TEST INFRASTRUCTURE ONLY).

Only usable in the authoring container: ``/root/reference`` does not exist on
the GPU box, so nothing in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may
call this.  It is used by ``oracle/make_golden.py`` (to freeze golden vectors
under ``tests/golden/``) and by ``tests/test_oracle_vs_reference.py`` (skipped
when the reference tree is absent).

The reference does not run at HEAD on this stack; three arithmetic-neutral
shims are applied:
  1. stub modules for hdbscan / matplotlib(.pyplot) / seaborn / umap
     (imported at kmer.py:6-10, unused by the profile path);
  2. ``numpy.float = float``  (kmer.py:207 uses the alias removed in NumPy 1.24);
  3. ``KmerClustering._KmerClustering__is_palindrome`` aliased to the public
     static ``is_palindrome`` (kmer.py:79/:194 call the name-mangled private
     name, but the method was renamed at kmer.py:46-47).
"""
import os
import sys
import tempfile
import types

REFERENCE_ROOT = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "karma", "kmer.py"))


_cls = None


def load():
    """Return the reference's KmerClustering class (shimmed)."""
    global _cls
    if _cls is not None:
        return _cls
    if not available():
        raise RuntimeError("reference tree not present at /root/reference")
    import numpy as np
    for name in ("hdbscan", "matplotlib", "matplotlib.pyplot", "seaborn", "umap"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if not hasattr(np, "float"):
        np.float = float
    # ``import karma.logs`` opens ./karma.log (logs.py:14): do it from a scratch dir
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="karma_ref_")
    os.chdir(scratch)
    # NOTE: the reference modules stay registered as ``karma`` / ``karma.kmer``:
    # its Pool.starmap (kmer.py:218) pickles a bound method, which needs the
    # class to be importable under its own module name.  This repo ships no
    # package called ``karma`` so nothing is shadowed.
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import karma.kmer as ref_kmer  # noqa: the reference module
        import karma.logs as ref_logs
        ref_logs.logger.setLevel(100)  # silence; does not change arithmetic
    finally:
        sys.path.remove(REFERENCE_ROOT)
        os.chdir(cwd)
    cls = ref_kmer.KmerClustering
    cls._KmerClustering__is_palindrome = staticmethod(cls.is_palindrome)
    _cls = cls
    return cls


def reference_profile(sequences, kmer_size="5p6", threads=2):
    """Run the reference's __calc_kmer_profile.  Returns (columns, matrix).
    Propagates SystemExit(1) exactly as the reference does."""
    cls = load()
    obj = cls(sequences, tempfile.gettempdir(), kmer_size, threads)
    mat = obj._KmerClustering__calc_kmer_profile()
    cols = sorted(obj.kmers, key=obj.kmers.get)
    return cols, mat
