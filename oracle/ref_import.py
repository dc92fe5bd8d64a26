"""oracle/ref_import.py -- import the byte-compiled reference under oracle/_ref (see make_ref.py).

TEST / BENCH INFRASTRUCTURE ONLY.  The compiled modules are stored as `name.refc` (CPython bytecode, exactly what a
`name.pyc` holds; the snapshot that carries the repo to the GPU box drops `*.pyc` files, so they cannot keep that
extension).  `install(roots)` puts a finder on sys.meta_path that resolves `import a.b` against those files the way the
path finder resolves it against `.py` / `.pyc`: `a/__init__.refc` is a package, `a.refc` a module."""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys

EXT = ".refc"


class _RefFinder(importlib.abc.MetaPathFinder):
    def __init__(self, roots):
        self.roots = list(roots)

    def find_spec(self, fullname, path=None, target=None):
        name = fullname.rpartition(".")[2]
        for root in (path if path is not None else self.roots):
            base = os.path.join(root, name)
            init = os.path.join(base, "__init__" + EXT)
            if os.path.isfile(init):
                loader = importlib.machinery.SourcelessFileLoader(fullname, init)
                return importlib.util.spec_from_file_location(fullname, init, loader=loader, submodule_search_locations=[base])
            if os.path.isfile(base + EXT):
                loader = importlib.machinery.SourcelessFileLoader(fullname, base + EXT)
                return importlib.util.spec_from_file_location(fullname, base + EXT, loader=loader)
        return None


def install(roots):
    sys.meta_path.append(_RefFinder(roots))
