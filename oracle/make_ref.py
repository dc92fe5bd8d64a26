#!/usr/bin/env python
"""oracle/make_ref.py -- byte-compile the UNMODIFIED reference into oracle/_ref/ so that it can be timed on the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY (the cpu_baseline leg and `bench.py --impl reference`).  /root/reference does not
exist on the GPU box; the reference is pure Python, so its "build" is CPython's own compiler: every .py under
/root/reference/src is compiled where it lies (py_compile, no source is copied) and only the bytecode OUTPUT is written
under oracle/_ref/ (git-ignored, not gpurun-ignored: it travels with the snapshot like our own .so files) as
`name.refc` -- the content of a `name.pyc` under an extension the snapshot does not filter out.  oracle/ref_import.py
makes oracle/_ref/src importable exactly like /root/reference/src, behind the same import shims (oracle/stubs: base
classes + registry only, no game logic).

    python oracle/make_ref.py        # no-op (exit 0) when /root/reference is absent
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")


def make_ref(verbose=True):
    import shutil
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        if verbose:
            print("make_ref: /root/reference/src is absent; keeping whatever oracle/_ref holds")
        return False
    shutil.rmtree(OUT, ignore_errors=True)
    n = 0
    for dirpath, dirnames, filenames in os.walk(src):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, REF)
        for fn in filenames:
            if not fn.endswith(".py"):
                continue
            dst = os.path.join(OUT, rel, fn[:-3] + ".refc")
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            # dfile: the path shown in tracebacks is the reference's own
            py_compile.compile(os.path.join(dirpath, fn), cfile=dst, dfile=os.path.join("/root/reference", rel, fn),
                               doraise=True, optimize=0, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            n += 1
    with open(os.path.join(OUT, "README"), "w") as f:
        f.write("CPython %d.%d bytecode of /root/reference/src, written by oracle/make_ref.py; not source, not tracked\n"
                % sys.version_info[:2])
    if verbose:
        print(f"make_ref: compiled {n} modules into {OUT}")
    return True


if __name__ == "__main__":
    make_ref()
