"""Build the oracle's native pieces (TEST INFRASTRUCTURE — see oracle/oracle_c.c header).

  oracle/_build/liboracle_c.so   gcc -O2 -ffp-contract=off oracle_c.c        (always)
  oracle/_ref/index_max*.so  the REFERENCE's own plugin, compiled from the sources where they
                                 lie under /root/reference/models/index_max_ext (read-only, never
                                 copied), with torch.utils.cpp_extension — only when /root/reference
                                 exists (this container). The GPU box uses the prebuilt file.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REFDIR = os.path.join(HERE, "_ref")
LIB = os.path.join(BUILD, "liboracle_c.so")
REF_SRC = "/root/reference/models/index_max_ext"


def build_c(force=False):
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "oracle_c.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(src) > os.path.getmtime(LIB):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", LIB,
                               src, "-lm"])
        print("[oracle] built", os.path.relpath(LIB, HERE))
    return LIB


def ref_plugin_path():
    hits = sorted(glob.glob(os.path.join(REFDIR, "index_max*.so")))
    return hits[0] if hits else None


def build_ref(force=False):
    """Compile the reference plugin in place (sources are only read). ~2-3 minutes, once."""
    if ref_plugin_path() and not force:
        return ref_plugin_path()
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(REFDIR, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    scratch = os.path.join(REFDIR, "_scratch")
    os.makedirs(scratch, exist_ok=True)
    load(name="index_max",
         sources=[os.path.join(REF_SRC, "index_max.cpp"), os.path.join(REF_SRC, "index_max_cuda.cu")],
         build_directory=scratch, verbose=False)
    built = glob.glob(os.path.join(scratch, "index_max*.so"))
    for b in built:
        os.replace(b, os.path.join(REFDIR, os.path.basename(b)))
    # keep only the shared object (objects/ninja files are scratch)
    for f in glob.glob(os.path.join(scratch, "*")):
        try:
            os.remove(f)
        except OSError:
            pass
    try:
        os.rmdir(scratch)
    except OSError:
        pass
    print("[oracle] built reference plugin ->", ref_plugin_path())
    return ref_plugin_path()


if __name__ == "__main__":
    build_c(force="--force" in sys.argv)
    if "--no-ref" not in sys.argv:
        build_ref(force="--force-ref" in sys.argv)
